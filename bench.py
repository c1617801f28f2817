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
is a Python package that cannot travel to the GPU box), bounded sample, same metric/unit.
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="A", choices=list(SHAPES))
    ap.add_argument("--batch", type=int, default=4, help="prompts per GPU per step")
    ap.add_argument("--precision", default="tf32", choices=["tf32", "fp16", "bf16", "fp32"])
    ap.add_argument("--mp-steps", type=int, default=20)
    ap.add_argument("--vae-precision", default=None, choices=["tf32", "fp32"],
                    help="VQGAN conv precision (default: tf32 tensor cores unless --precision fp32)")
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
def cpu_reference_tokens_per_s(shape, mp_steps, n_fwd=2):
    """Times the oracle port of BERT.generate_images on the host cores on a BOUNDED sample:
    n_fwd of the mp_steps transformer+head forwards of one sample and 1 of the 8 frame decodes, extrapolated
    linearly (every mask-predict step is the same full forward; every frame decode is identical work)."""
    from mmvid_b200 import synth
    from oracle import mmvid_oracle as O
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    cfg = SHAPES[shape]
    spec = O.BertSpec(dim=DIM, text_seq_len=cfg["text_seq_len"], num_text_tokens=VOCAB, num_visuals=0,
                      num_targets=cfg["num_targets"], image_size=cfg["image_size"])
    g = torch.Generator().manual_seed(0)

    def rnd(*s, scale=0.02):
        return torch.randn(*s, generator=g) * scale

    sd = {}
    for k, shp in synth.resblock_keys("transformer.transformer.", DIM, LAYERS).items():
        sd[k] = torch.ones(shp) if (k.endswith("weight") and len(shp) == 1) else rnd(*shp)
    sd["image_emb.weight"] = rnd(1026, DIM, scale=1.0)
    for i, s in enumerate(((1, cfg["num_targets"], 1, 1, DIM), (1, 1, spec.fmap, 1, DIM), (1, 1, 1, spec.fmap, DIM))):
        sd[f"target_pos_emb.weights_{i}"] = rnd(*s, scale=1.0)
    sd["to_logits.0.weight"], sd["to_logits.0.bias"] = torch.ones(DIM), torch.zeros(DIM)
    sd["to_logits.1.weight"], sd["to_logits.1.bias"] = rnd(1024, DIM), torch.zeros(1024)
    control = rnd(1, spec.control_seq_len, DIM, scale=1.0)
    tgt = torch.full((1, spec.target_seq_len), spec.MASK, dtype=torch.long)
    with torch.no_grad():
        O.bert_logits(spec, sd, control, tgt)  # warm-up
        ts = []
        for _ in range(n_fwd):
            t0 = time.perf_counter()
            logits = O.bert_logits(spec, sd, control, tgt)
            probs = torch.softmax(logits, -1)
            torch.multinomial(probs[0], 1)
            ts.append(time.perf_counter() - t0)
        t_fwd = statistics.median(ts)
        # VQGAN decode of one frame with default-initialised weights
        from mmvid_b200.vae import VQGanVAE1024
        vae = VQGanVAE1024(image_size=cfg["image_size"])
        vsd = {k: v.detach() for k, v in vae.state_dict().items()}
        ids = torch.randint(0, 1024, (1, spec.image_seq_len), generator=g)
        O.vae_decode(ids, vsd)
        t0 = time.perf_counter()
        O.vae_decode(ids, vsd)
        t_dec = time.perf_counter() - t0
    t_sample = mp_steps * t_fwd + cfg["num_targets"] * t_dec
    return dict(value=spec.target_seq_len / t_sample, unit="video-tokens/s", cores=cores, kind="port",
                sample=f"{n_fwd} of {mp_steps} transformer+head forwards (S={spec.total_seq_len}) of 1 prompt and 1 of "
                       f"{cfg['num_targets']} frame decodes, fp32 torch CPU, extrapolated linearly",
                t_forward_s=round(t_fwd, 3), t_decode_frame_s=round(t_dec, 3))


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps_vals = []
    info = None
    for _ in range(max(1, min(args.steps, 2))):
        info = cpu_reference_tokens_per_s(args.shape, args.mp_steps, n_fwd=1)
        steps_vals.append(info["value"])
    v = statistics.median(steps_vals)
    cfg = SHAPES[args.shape]
    out = {
        "impl": "reference", "metric": "video-tokens/sec (BERT.generate_images, mask-predict T=%d + VQGAN decode)" % args.mp_steps,
        "value": v, "unit": "video-tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * cfg["num_targets"] * (cfg["image_size"] // 16) ** 2 / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": dict(info, value=v),
        "e2e": {"value": v, "unit": "video-tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)


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
            "precision": args.precision,
            "vae_precision": args.vae_precision or ("fp32" if args.precision == "fp32" else "tf32"),
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


def build_artv(args, device):
    from mmvid_b200.dalle_artv import DALLE
    from mmvid_b200.vae import VQGanVAE1024
    cfg = SHAPES[args.shape]
    torch.manual_seed(1234)
    vprec = args.vae_precision or ("fp32" if args.precision == "fp32" else "tf32")
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
    vprec = args.vae_precision or ("fp32" if args.precision == "fp32" else "tf32")
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
    return out


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
    host_frames = torch.empty(B * world, cfg["num_targets"], 3, cfg["image_size"], cfg["image_size"]).pin_memory()
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
    d2h_bytes = host_frames.numel() * 4
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
        if world > 1:
            images = all_gather_variable(images.contiguous(), [B] * world)
        if e2e:
            host_frames.copy_(images, non_blocking=True)
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
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
            if dom == "attention" and args.precision == "tf32" and args.shape == "A" and B == 4:
                traffic = tj["attention_tc3_kernel_tf32_shapeA_b4"]["dram_bytes_per_launch"]
        except Exception:
            traffic = None
        roof = {"bound": "tensor", "kernel": dom, "achieved": kr[dom]["tflops"], "peak": peak, "unit": "TFLOP/s",
                "frac": kr[dom]["tflops"] / peak, "traffic": traffic, "peak_source": peaks["source"] + " cuBLAS bf16 burst",
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
        if not args.no_cpu_baseline and world == 1 and args.workload == "bert":
            out["cpu_baseline"] = cpu_reference_tokens_per_s(args.shape, args.mp_steps, n_fwd=2)
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
