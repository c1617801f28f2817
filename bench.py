"""Headline benchmark: end-to-end video-tokens/sec of BERT.generate_images (mask-predict + VQGAN decode).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shape A|B] [--batch B]
                    [--precision tf32|bf16|fp32] [--mp-steps T]

Workload (BASELINE.json `metric`): text-to-video, Shape A = text 64 + 8 frames x (16 x 16) VQGAN
tokens (S = 2115, image_size 256 because the shipped VQGAN is f16), CLIP ViT-B/32-shaped transformer (768 wide,
12 layers, 12 heads, random init), mask-predict T = 20, beam 1, static schedule, then VQGAN decode of the 8
frames.  One "step" = one generate_images call on a batch of `--batch` prompts per GPU.
   value  video tokens finalised per second with the prompt already in HBM (whole job, all ranks);
   e2e    same through the public API with host buffers: pinned H2D of the token ids, D2H of the frames.
Multi-GPU: replicas, the sampling batch is split across ranks (weak scaling: per-GPU batch fixed), one NCCL
all-gather of the decoded frames per step (inside the timed region).
`--impl reference`: the reference's own algorithm on the host CPU cores (oracle port; the reference itself
is a Python package that cannot travel to the GPU box): every step is ONE FULL prompt - control embedding, all T
mask-predict iterations (forward, head, sampling, keep-mask multinomial) and the decode of all frames - of the same
configuration; the reference loops over the prompts of a batch serially (dalle_bert.py:618), so its tokens/s does not
depend on the batch.  The number of executed steps is capped so that the arm ends within a few minutes.
`--impl eager`: the same algorithm in plain PyTorch eager on the GPU (oracle restatement on CUDA, TF32 allowed, torch
SDPA): the library baseline BASELINE.md section 3 asks for.  Not the product path: nothing of mmvid_b200 runs in it.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SHAPES = {
    "A": dict(text_seq_len=64, image_size=256, num_targets=8),   # seq 2115: the metric BASELINE.json names
    "B": dict(text_seq_len=50, image_size=128, num_targets=8),   # seq 565: what the reference's scripts run
}
VOCAB = 49408
DIM, LAYERS = 768, 12


_REAL_STDOUT = None


def quiet_stdout():
    """From here on everything written to file descriptor 1 - Python prints AND C libraries such as NCCL, which prints its
    version banner there - goes to stderr; the one JSON line is written to the saved descriptor by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, line)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "eager"])
    ap.add_argument("--shape", default="A", choices=list(SHAPES))
    ap.add_argument("--batch", type=int, default=4, help="prompts per GPU per step")
    ap.add_argument("--precision", default="fp16", choices=["tf32", "fp16", "bf16", "fp32"],
                    help="fp16 (default): kind::f16 with fp16 operands = the tf32 mantissa at twice the rate; logits / pixels stay "
                         "within the 1e-3 bar of north_star (tests/test_gpu_0_models.py); bf16 does not (2.7e-3)")
    ap.add_argument("--mp-steps", type=int, default=20)
    ap.add_argument("--vae-precision", default=None, choices=["fp16", "tf32", "fp32"],
                    help="VQGAN decoder conv precision (default: fp16 with --precision fp16, fp32 with fp32, else tf32)")
    ap.add_argument("--visuals", type=int, default=0, choices=[0, 1],
                    help="bert/train: number of visual-control frames (1 = SURVEY config 4: cVAE-encoded frame, S=2371)")
    ap.add_argument("--workload", default="bert", choices=["bert", "artv", "train"],
                    help="bert: BERT mask-predict generate_images (headline); artv: ART-V KV-cache autoregressive generate_images")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true", help="ncu mode: no warm-up, one step, no side measurements")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d.get("bf16_tflops_sustained"),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# --------------------------------------------------------------------------------------------------- CPU arm
def oracle_problem(args, device):
    """(spec, state dict with the reference's key names, text [1, L], visual or None) of ONE prompt of the benchmark
    configuration, random-init weights of the benchmark architecture (the same constructor the GPU arm uses)."""
    from oracle import mmvid_oracle as O
    cfg = SHAPES[args.shape]
    V = getattr(args, "visuals", 0)
    model = build_model(argparse.Namespace(**dict(vars(args), workload="bert", precision="fp32")), torch.device("cpu"))
    sd = {k: v.detach().to(device) for k, v in model.state_dict().items()}
    spec = O.BertSpec(dim=DIM, text_seq_len=cfg["text_seq_len"], num_text_tokens=VOCAB, num_visuals=V,
                      num_targets=cfg["num_targets"], image_size=cfg["image_size"], has_cvae=V > 0)
    g = torch.Generator().manual_seed(42)
    text = torch.randint(1, VOCAB, (1, cfg["text_seq_len"]), generator=g)
    text[:, -cfg["text_seq_len"] // 4:] = 0
    visual = torch.rand(1, 1, 3, cfg["image_size"], cfg["image_size"], generator=g) if V else None
    return spec, sd, text.to(device), (visual.to(device) if visual is not None else None)


def cpu_full_prompts(args, budget_s, max_steps, warm=True):
    """Times FULL prompts of the reference algorithm (oracle port) on all host cores: per prompt the control embedding
    (+ cVAE encode of the visual control when the configuration has one), all T mask-predict iterations and the decode
    of all frames.  Executes at least one prompt, then as many more (up to max_steps) as fit into budget_s."""
    from oracle import mmvid_oracle as O
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    spec, sd, text, visual = oracle_problem(args, torch.device("cpu"))
    mpc = dict(O.DEFAULT_MP_CONFIG, T=args.mp_steps)
    with torch.no_grad():
        if warm:  # one untimed transformer forward: thread pool, oneDNN primitive caches, page faults of the weights
            c = O.bert_control_emb(spec, sd, text, None)
            O.bert_logits(spec, sd, c, torch.full((1, spec.target_seq_len), spec.MASK, dtype=torch.long))
        times = []
        t_start = time.perf_counter()
        while True:
            torch.manual_seed(42 + len(times))
            t0 = time.perf_counter()
            images, seq = O.bert_generate_images(spec, sd, text, visual, steps=args.mp_steps, mp_config=mpc, dynamic=False)
            times.append(time.perf_counter() - t0)
            assert images.shape[1] == spec.num_targets and seq.numel() == spec.target_seq_len
            elapsed = time.perf_counter() - t_start
            if len(times) >= max_steps or elapsed + statistics.mean(times) > budget_s:
                break
    tps = spec.target_seq_len * len(times) / sum(times)
    return dict(value=tps, unit="video-tokens/s", cores=cores, kind="port",
                sample=f"{len(times)} full prompt(s) of the configuration, each = control embedding + {args.mp_steps} mask-predict "
                       f"iterations (S={spec.total_seq_len} forward, head, sampling, keep-mask multinomial) + "
                       f"{spec.num_targets} frame decodes; fp32 torch on {cores} host threads; the reference processes the "
                       "prompts of a batch one after the other (dalle_bert.py:618), so tokens/s is batch-independent",
                s_per_prompt=[round(t, 3) for t in times]), times, spec


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload != "bert":
        emit({"impl": "reference", "unavailable": "the CPU arm times the headline workload (BERT.generate_images) only"})
        return
    info, times, spec = cpu_full_prompts(args, budget_s=150.0, max_steps=max(1, args.steps))
    out = {
        "impl": "reference", "metric": "video-tokens/sec (BERT.generate_images, mask-predict T=%d + VQGAN decode)" % args.mp_steps,
        "value": info["value"], "unit": "video-tokens/s", "n_gpus": args.gpus, "steps": len(times), "warmup": 1,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": 1000.0 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.batch),
        "step": "one full prompt of the configuration's batch (the reference's sample loop is serial); warm-up = one "
                "untimed transformer forward; executed steps capped so that the arm ends within ~150 s",
        "cpu_baseline": info,
        "e2e": {"value": info["value"], "unit": "video-tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


def run_eager_arm(args):
    """PyTorch-eager fp32-storage GPU baseline: the oracle restatement of the reference on CUDA
    with TF32 matmuls allowed and torch's SDPA, sample-serial like the reference.  Library kernels only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    assert torch.cuda.is_available()
    from oracle import mmvid_oracle as O
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    spec, sd, text, visual = oracle_problem(args, dev)
    B = args.batch
    text = text.repeat(B, 1)
    visual = visual.repeat(B, 1, 1, 1, 1) if visual is not None else None
    mpc = dict(O.DEFAULT_MP_CONFIG, T=args.mp_steps)

    def step():
        return O.bert_generate_images(spec, sd, text, visual, steps=args.mp_steps, mp_config=mpc, dynamic=False)

    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 3))):
            step()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(args.steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
    v = spec.target_seq_len * B * args.steps / (tot / 1000.0)
    emit({"impl": "eager", "metric": "video-tokens/sec (BERT.generate_images, mask-predict T=%d + VQGAN decode)" % args.mp_steps,
          "value": v, "unit": "video-tokens/s", "n_gpus": 1, "steps": args.steps, "warmup": max(1, min(args.warmup, 3)),
          "ms_per_step": tot / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
          "dtype": "tf32 (torch eager, allow_tf32=True, SDPA)", "data": "synthetic", "config": workload_config(args, B),
          "note": "library baseline: torch / cuBLAS / cuDNN kernels running the oracle restatement of the reference on the GPU, "
                  "sample-serial like the reference (dalle_bert.py:618); nothing of mmvid_b200 is on this path",
          "gpu_launches": 0})


def workload_config(args, per_gpu_batch):
    cfg = SHAPES[args.shape]
    fmap = cfg["image_size"] // 16
    V = getattr(args, "visuals", 0)
    S = 1 + cfg["text_seq_len"] + V * fmap * fmap + 2 + cfg["num_targets"] * fmap * fmap
    name = (f"BERT.generate_images text-to-video shape {args.shape}: text {cfg['text_seq_len']} + "
            f"{cfg['num_targets']}x({fmap}x{fmap}) video tokens (S={S}), ViT-B/32-shaped transformer 768x12, "
            f"mask-predict T={args.mp_steps} beam 1 + VQGAN f16 decode @{cfg['image_size']}px")
    if V:
        name += f", {V} visual-control frame(s) through the cVAE encoder"
    if getattr(args, "workload", "bert") == "train":
        name = (f"BERT.forward(return_loss=True, rel=True, vid=True) + backward + clip_grad_norm_ + Adam, shape {args.shape} "
                f"(S={S}): 2 VQGAN encodes of the {cfg['num_targets']} target frames, 3 transformer passes (train.py:298-325)")
    return {"workload": name,
            "per_gpu_batch": per_gpu_batch, "global_batch": per_gpu_batch * args.gpus, "seq_len": S,
            "parallelism": f"replicas x{args.gpus}, batch split, one all-gather of frames",
            "l2": "explicit 256 MiB L2 flush between timed steps (outside the timed events)"}


# --------------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={dev}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], None, set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def vae_precision(args):
    return args.vae_precision or {"fp32": "fp32", "fp16": "fp16"}.get(args.precision, "tf32")


def build_artv(args, device):
    from mmvid_b200.dalle_artv import DALLE
    from mmvid_b200.vae import VQGanVAE1024
    cfg = SHAPES[args.shape]
    torch.manual_seed(1234)
    vprec = vae_precision(args)
    vae = VQGanVAE1024(vae_path=None, image_size=cfg["image_size"], precision=vprec)
    vae.image_size = cfg["image_size"]
    vae.model.quantize.embedding.weight.data.normal_(0, 0.3)
    model = DALLE(dim=DIM, vae=vae, cvae=vae, num_text_tokens=VOCAB, text_seq_len=cfg["text_seq_len"],
                  which_transformer="openai_clip_visual", num_visuals=1, num_targets=cfg["num_targets"],
                  openai_clip_path=None, transformer_layers=LAYERS, precision=args.precision, sampling_mode="batched")
    return model.to(device).eval()


def build_model(args, device):
    if args.workload == "artv":
        return build_artv(args, device)
    if args.workload == "train" and args.precision in ("bf16", "fp16"):
        raise SystemExit("training runs the tf32 or fp32 path (16-bit activations are an inference-only mode)")
    from mmvid_b200.dalle_bert import BERT
    from mmvid_b200.vae import VQGanVAE1024
    cfg = SHAPES[args.shape]
    torch.manual_seed(1234)
    vprec = vae_precision(args)
    vae = VQGanVAE1024(vae_path=None, image_size=cfg["image_size"], precision=vprec)
    vae.image_size = cfg["image_size"]
    # default VQ init U(+-1/1024) is degenerate for decoding; use unit-scale codes (random-init weights, no checkpoint)
    vae.model.quantize.embedding.weight.data.normal_(0, 0.3)
    cvae = None
    if args.visuals:
        cvae = VQGanVAE1024(vae_path=None, image_size=cfg["image_size"], precision=vprec)
        cvae.image_size = cfg["image_size"]
        cvae.model.quantize.embedding.weight.data.normal_(0, 0.3)
    model = BERT(dim=DIM, vae=vae, cvae=cvae, num_text_tokens=VOCAB, text_seq_len=cfg["text_seq_len"],
                 which_transformer="openai_clip_visual", num_visuals=args.visuals, num_targets=cfg["num_targets"],
                 openai_clip_path=None, transformer_layers=LAYERS, precision=args.precision, sampling_mode="batched")
    model = model.to(device)
    return model.train() if args.workload == "train" else model.eval()


def kernel_roofline(model, args, peaks, S, B):
    """Times the dominant kernels alone at the benchmark shapes with CUDA events on the launching stream."""
    from mmvid_b200 import ops
    from mmvid_b200._lib import PRECISIONS
    prec = PRECISIONS[args.precision]
    dev = next(model.parameters()).device
    H = DIM // 64
    act_dt = ops.act_dtype(prec)
    qkv = torch.randn(B * S, 3 * DIM, device=dev)
    x = torch.randn(B * S, DIM, device=dev).to(act_dt)
    blk = model.transformer.transformer.resblocks[0]
    w_fc = model.transformer._w(blk.mlp.c_fc.weight, prec)
    flush = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32)

    def timeit(fn, reps=10):
        fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot / reps  # ms

    out = {}
    if prec != 0:
        # attention core alone: split + flash kernel; FLOPs = 4 S^2 D per batch element (QK^T + PV)
        S_pad = (S + 127) // 128 * 128
        import ctypes as C
        from mmvid_b200 import _lib as L
        lib = L.load()
        dt = act_dt
        q = torch.empty(B, H, S_pad, 64, device=dev, dtype=dt)
        k = torch.empty_like(q)
        vt = torch.empty(B, H, 64, S_pad, device=dev, dtype=dt)
        L.check(lib.mmvid_qkv_split(ops._ptr(qkv), ops._ptr(q), ops._ptr(k), ops._ptr(vt), ops._dt(q), B, H, S, S_pad,
                                    ops._stream()))
        o = torch.empty(B * S, DIM, device=dev, dtype=act_dt)
        pr = (C.c_int * 4)(*model.transformer.mask_rows, 0, 0)

        def att():
            L.check(lib.mmvid_attention(ops._ptr(q), ops._ptr(k), ops._ptr(vt), ops._ptr(o), ops._dt(o), o.stride(0), B, H,
                                        S, S_pad, model.transformer.mask_kind, pr, len(model.transformer.mask_rows), prec,
                                        ops._stream()))
        ms = timeit(att)
        fl = 4.0 * S * S * DIM * B
        out["attention"] = dict(ms=ms, tflops=fl / ms / 1e9)
    # MLP c_fc GEMM: [B*S,768] x [768,3072]
    def fc():
        ops.linear(x, w_fc, blk.mlp.c_fc.bias, act=1, precision=prec, out_dtype=act_dt)
    ms = timeit(fc)
    out["gemm_c_fc"] = dict(ms=ms, tflops=2.0 * B * S * DIM * 4 * DIM / ms / 1e9)
    # ---- HBM-bound kernels: algorithmic bytes (each operand read once, each result written once) / time, against the
    # measured copy bandwidth.  esz = bytes per element of the 16-bit / fp32 activation the kernel writes.
    hbm = peaks["hbm_gbs"]
    esz = 2 if act_dt != torch.float32 else 4

    def gbs(nbytes, ms_):
        return dict(ms=ms_, gbs=nbytes / ms_ / 1e6, frac_of_hbm=nbytes / ms_ / 1e6 / hbm, bytes=nbytes)

    xr = torch.randn(B * S, DIM, device=dev)
    ms = timeit(lambda: ops.layernorm(xr, blk.ln_1.weight, blk.ln_1.bias, 1e-5, out_dtype=act_dt))
    out["layernorm"] = gbs(B * S * DIM * (4 + esz), ms)
    cfg = SHAPES[args.shape]
    fmap = cfg["image_size"] // 16
    Ttot = cfg["num_targets"] * fmap * fmap
    ids = torch.randint(0, 1024, (B, Ttot), device=dev)
    xg = torch.empty(B, model.total_seq_len, DIM, device=dev)
    seg = model._target_segment(ids)
    ms = timeit(lambda: ops.embed_gather(xg, [seg]))
    out["embed_gather"] = gbs(B * Ttot * DIM * 4 * 2 + Ttot * DIM * 4, ms)  # table rows in + rows out (+ position table once)
    z = torch.randn(B * Ttot, 256, device=dev)
    cb = model.vae.model.quantize.embedding.weight.detach()
    ms = timeit(lambda: ops.vq_argmin(z, cb))
    # distances are 2 x T x 1024 x 256 fp32 FLOP on the CUDA cores (bit-exact indices need fp32); bytes are negligible
    out["vq_argmin"] = dict(ms=ms, tflops_fp32=2.0 * B * Ttot * cb.shape[0] * 256 / ms / 1e9, bound="fp32 FFMA",
                            bytes=B * Ttot * (256 * 4 + 8) + cb.numel() * 4)
    gx = torch.randn(B * cfg["num_targets"], cfg["image_size"] // 2, cfg["image_size"] // 2, 128, device=dev)
    gw, gb_ = torch.ones(128, device=dev), torch.zeros(128, device=dev)
    ms = timeit(lambda: ops.groupnorm(gx, gw, gb_, swish=True, fast=True))
    out["groupnorm_swish"] = gbs(gx.numel() * 4 * 3, ms)  # statistics pass + apply pass read, one write (K11)
    return out


def decode_roofline(args, peaks, B, S_max, pos):
    """ART-V workload: the streaming decode kernel (decode_stream.cu) alone, one token of all B samples at cache length
    `pos` (the mean length over the generation), random 16-bit weights of the benchmark architecture.  HBM roofline:
    algorithmic bytes = every 16-bit weight once (12 layers x 12 D^2 + head) + the K/V cache rows read (2 x layers x B x
    pos x D x 2 B); the weights alone (170 MB) exceed the 126 MB L2, so back-to-back launches are L2-cold for them."""
    import ctypes as C
    from mmvid_b200 import _lib as L, ops
    lib = L.load()
    dev = "cuda"
    D, H, NL = DIM, DIM // 64, LAYERS
    f16 = 0 if args.precision == "bf16" else 1
    dt = torch.bfloat16 if args.precision == "bf16" else torch.float16
    g = torch.Generator().manual_seed(0)

    def r(*s):
        return (torch.randn(*s, generator=g) * 0.02).to(dev)
    layers16 = (L.DecodeLayer16 * NL)()
    keep = []
    for li in range(NL):
        t = dict(ln1_w=torch.ones(D, device=dev), ln1_b=torch.zeros(D, device=dev), in_b=r(3 * D), out_b=r(D),
                 ln2_w=torch.ones(D, device=dev), ln2_b=torch.zeros(D, device=dev), fc_b=r(4 * D), proj_b=r(D),
                 in_w=r(3 * D, D).to(dt), out_w=r(D, D).to(dt), fc_w=r(4 * D, D).to(dt), proj_w=r(D, 4 * D).to(dt),
                 kcache=r(B, H, S_max, 64).to(dt), vcache=r(B, H, S_max, 64).to(dt))
        keep.append(t)
        for k, v in t.items():
            setattr(layers16[li], k, v.data_ptr())
    ws = torch.zeros(int(lib.mmvid_artv_decode_stream_workspace_floats(B, D, H)), device=dev)
    head_w, head_b = r(1024, D).to(dt), r(1024)
    lnw, lnb = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    logits = torch.empty(B, 1024, device=dev)
    h = r(B, D) * 50
    st = ops._stream()

    def call():
        L.check(lib.mmvid_artv_decode_stream(layers16, NL, ops._ptr(h), ops._ptr(ws), ops._ptr(lnw), ops._ptr(lnb),
                                             ops._ptr(head_w), ops._ptr(head_b), ops._ptr(logits), 1024, B, D, H, S_max,
                                             pos, None, f16, st))
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    reps = 20
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        call()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    w_bytes = (NL * 12 * D * D + 1024 * D) * 2
    kv_bytes = 2 * NL * B * (pos + 1) * D * 2
    nbytes = w_bytes + kv_bytes
    gbs = nbytes / ms / 1e6
    return {"bound": "hbm", "kernel": "artv_decode_stream", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": gbs / peaks["hbm_gbs"], "traffic": None, "algorithmic_bytes": nbytes, "ms": ms, "cache_len": pos,
            "peak_source": peaks["source"] + " copy bandwidth",
            "note": "one launch = one token of all %d samples (12 layers + head); weights %.0f MB + K/V %.0f MB per launch; the "
                    "step is a chain of 61 grid-wide phases, so it is latency- not byte-bound (DESIGN 2.4)" % (B, w_bytes / 1e6, kv_bytes / 1e6)}


def parity_selfcheck(model, args, S, dev):
    """Outside the timed region: logits of ONE forward (1 prompt, random masked targets) in the benchmarked precision against
    the library's own fp32 CUDA-core path (itself checked against the reference at 7e-7 in tests/).  The <= 1e-3 bar of
    north_star is asserted by the GPU tests on the reference's golden logits; this puts the number next to the headline."""
    from mmvid_b200 import ops
    if args.precision == "fp32" or args.workload != "bert":
        return None
    cfg = SHAPES[args.shape]
    g = torch.Generator().manual_seed(7)
    text = torch.randint(1, VOCAB, (1, cfg["text_seq_len"]), generator=g).to(dev)
    tgt = torch.randint(0, 1024, (1, model.target_seq_len), generator=g)
    tgt[:, ::3] = 1024
    tgt = tgt.to(dev)
    res = {}
    keep = (model.precision, model.transformer.precision)
    for prec in ("fp32", args.precision):
        model.precision = model.transformer.precision = prec
        control = model(text, visual=None, return_loss=False)
        x = torch.empty(1, model.total_seq_len, DIM, device=dev)
        x[:, :control.shape[1]] = control
        ops.embed_gather(x, [model._target_segment(tgt)])
        hid = model.transformer_forward(x)
        res[prec] = model._head(hid[:, control.shape[1]:].reshape(-1, DIM), model.to_logits).double()
    model.precision, model.transformer.precision = keep
    a, b = res[args.precision], res["fp32"]
    return {"logits_relerr_vs_own_fp32_path": float((a - b).norm() / b.norm()),
            "argmax_agreement": float((a.argmax(-1) == b.argmax(-1)).double().mean()),
            "what": "one Shape-%s forward, 1 prompt, 2/3 of the target tokens given; sampled ids in the benchmarked 'batched' "
                    "mode follow the reference's distribution but not its RNG order (bit-exact ids: fp32 mode + "
                    "sampling_mode='reference', tests/test_gpu_0_models.py)" % args.shape}


def run_ours(args):
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line (NCCL prints its version banner there)
        dist.init_process_group("nccl", device_id=dev)
    from mmvid_b200 import _lib
    from mmvid_b200.parallel import all_gather_variable, all_reduce_gradients
    peaks = load_peaks()
    cfg = SHAPES[args.shape]
    model = build_model(args, dev)
    B = args.batch
    fmap = cfg["image_size"] // 16
    tokens_per_sample = cfg["num_targets"] * fmap * fmap
    g = torch.Generator().manual_seed(42 + rank)
    host_text = torch.randint(1, VOCAB, (B, cfg["text_seq_len"]), generator=g)
    host_text[:, -cfg["text_seq_len"] // 4:] = 0
    host_text = host_text.pin_memory()
    # e2e: every rank moves ITS OWN frames to pinned host memory (a single rank copying the whole gathered batch serialises
    # world x 25 MB on one PCIe link: round 1 measured 0.89 e2e efficiency at 8 GPUs for exactly that reason); the all-gather
    # over NVLink still runs every step because the gathered tensor is the API's result
    host_frames = torch.empty(B, cfg["num_targets"], 3, cfg["image_size"], cfg["image_size"]).pin_memory()
    dev_text = host_text.to(dev)
    flush = torch.empty(64 * 1024 * 1024, device=dev, dtype=torch.float32)
    torch.manual_seed(42 + rank)  # train.py:87 seeds seed+rank the same way

    artv_visual = None
    if args.workload == "artv":
        artv_visual = torch.rand(B, 1, 3, cfg["image_size"], cfg["image_size"], generator=g).to(dev)
    host_visual = dev_visual = None
    if args.visuals and args.workload != "artv":
        host_visual = torch.rand(B, 1, 3, cfg["image_size"], cfg["image_size"], generator=g).pin_memory()
        dev_visual = host_visual.to(dev)
    h2d_bytes = host_text.numel() * 8 + (host_visual.numel() * 4 if host_visual is not None else 0)
    d2h_bytes = host_frames.numel() * 4 * world  # whole job: every rank copies its shard
    train = args.workload == "train"
    if train:
        # train.py:298-325 with the CLI defaults (utils_args.py:357-410): Adam lr 1e-4, clip 1.0, betas 7 / 0.5 / 0.5
        import random
        import numpy as np
        from mmvid_b200 import optim as FO
        np.random.seed(42 + rank)
        random.seed(42 + rank)
        params = [p for p in model.parameters() if p.requires_grad]
        opt = FO.FusedAdam(params, lr=1e-4, weight_decay=0.0)
        host_target = torch.rand(B, cfg["num_targets"], 3, cfg["image_size"], cfg["image_size"], generator=g).pin_memory()
        dev_target = host_target.to(dev)
        host_loss = torch.empty(1).pin_memory()
        h2d_bytes += host_target.numel() * 4
        d2h_bytes = 4

    def step(e2e):
        if e2e:
            text = host_text.to(dev, non_blocking=True)
            visual = host_visual.to(dev, non_blocking=True) if host_visual is not None else None
        else:
            text, visual = dev_text, dev_visual
        if train:
            target = host_target.to(dev, non_blocking=True) if e2e else dev_target
            l_msm, l_rel, l_vid = model(text, visual=visual, target=target, return_loss=True, rel=True, vid=True)
            loss = 7.0 * l_msm + 0.5 * l_rel + 0.5 * l_vid
            opt.zero_grad()
            loss.backward()
            if world > 1:
                all_reduce_gradients(params)  # DDP-equivalent gradient averaging (train.py:32), NCCL over NVLink
            FO.clip_grad_norm_(params, 1.0)
            opt.step()
            if e2e:
                host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
            return loss
        if args.workload == "artv":
            images, _, seq = model.generate_images(text, visual=artv_visual)
        else:
            images, _, seq = model.generate_images(text, visual=visual, mask_predict_steps=args.mp_steps, dynamic=False)
        local = images
        if world > 1:
            images = all_gather_variable(images.contiguous(), [B] * world)
        if e2e:
            host_frames.copy_(local, non_blocking=True)
        return images

    def timed(e2e, steps):
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step(e2e)
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        t = torch.tensor([tot], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t) / 1000.0  # seconds, max over ranks

    if args.profile:
        os.environ["MMVID_CUDA_GRAPH"] = "0"  # launch list = exactly one step (a graph would add its warm-up forwards)
        step(False)
        torch.cuda.synchronize()
        if world > 1:
            dist.destroy_process_group()
        return
    for _ in range(max(args.warmup, 3)):
        step(False)
    torch.cuda.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    _lib.reset_launch_count()
    t_dev = timed(False, args.steps)
    launches = _lib.launch_count()
    t_e2e = timed(True, args.steps)
    clocks = sampler.stop() if sampler else None
    total_tokens = tokens_per_sample * B * world * args.steps
    value, e2e_value = total_tokens / t_dev, total_tokens / t_e2e
    if rank == 0:
        S = 1 + cfg["text_seq_len"] + args.visuals * fmap * fmap + 2 + tokens_per_sample
        kr = kernel_roofline(model, args, peaks, S, B) if args.workload in ("bert", "train") else {"gemm_c_fc": dict(ms=0.0, tflops=0.0)}
        dom = "attention" if "attention" in kr else "gemm_c_fc"
        peak = peaks["bf16_tflops"]
        # DRAM bytes of one launch of the dominant kernel come from an `ncu --set full` capture of this same command
        # (bench.py cannot read hardware counters itself); profiles/ncu_traffic.json records them per (kernel, precision,
        # shape, batch) together with the capture they came from.  No matching capture => null.
        traffic, traffic_src = None, None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            ent = tj.get(f"{dom}_{args.precision}_shape{args.shape}_b{B}")
            if ent:
                traffic, traffic_src = ent["dram_bytes_per_launch"], ent["source"]
        except Exception:
            pass
        if args.workload == "artv" and args.precision != "fp32" and B <= 8:
            prefix = 1 + cfg["text_seq_len"] + fmap * fmap
            roof_artv = decode_roofline(args, peaks, B, prefix + tokens_per_sample, prefix + tokens_per_sample // 2)
        roof = {"bound": "tensor", "kernel": dom, "achieved": kr[dom]["tflops"], "peak": peak, "unit": "TFLOP/s",
                "frac": kr[dom]["tflops"] / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes": (3 * B * (DIM // 64) * S * 64 + B * S * DIM) * (2 if args.precision in ("bf16", "fp16") else 4)
                if dom == "attention" else None,
                "algorithmic_flop": 4.0 * S * S * DIM * B if dom == "attention" else None,
                "peak_source": peaks["source"] + " cuBLAS bf16 burst",
                "note": "kind::tf32 issues at half the kind::f16 rate; frac_of_half_rate = achieved / (peak/2)"
                        if args.precision == "tf32" else "",
                "frac_of_half_rate": kr[dom]["tflops"] / (peak / 2) if args.precision == "tf32" else None,
                "kernels": kr}
        out = {
            "metric": {"bert": "video-tokens/sec (BERT.generate_images, mask-predict T=%d + VQGAN decode)" % args.mp_steps,
                       "artv": "video-tokens/sec (DALLE.generate_images, ART-V, KV-cache decode + VQGAN decode)",
                       "train": "video-tokens/sec (BERT training step: forward + backward + clip + Adam)"}[args.workload],
            "value": value, "unit": "video-tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1000.0 * t_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"tf32": "tf32 (fp32 storage, fp32 accumulate)", "bf16": "bf16 (fp32 accumulate, fp32 residual stream)",
                      "fp16": "fp16 (fp32 accumulate, fp32 residual stream)", "fp32": "f32"}[args.precision],
            "data": "synthetic", "config": dict(workload_config(args, B), **({"workload": "DALLE(ART-V).generate_images with KV cache, shape " + args.shape + ": prefix 1+L+n, " + str(tokens_per_sample) + " decode steps, batch " + str(B)} if args.workload == "artv" else {})),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "video-tokens/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h_bytes, "ms_per_step": 1000.0 * t_e2e / args.steps},
            "gpu_launches": launches, "roofline": roof,
        }
        if args.workload == "artv" and args.precision != "fp32" and B <= 8:
            out["roofline"] = roof_artv
        out["parity"] = parity_selfcheck(model, args, S, dev)
        out["precision"] = {"transformer": args.precision,
                            "vae_decoder": vae_precision(args)}
        if not args.no_cpu_baseline and world == 1 and args.workload == "bert":
            out["cpu_baseline"] = cpu_full_prompts(args, budget_s=30.0, max_steps=1)[0]
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.impl == "eager":
        run_eager_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
