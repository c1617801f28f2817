"""GPU: size-independent properties at BASELINE.json's FULL sizes (Shape A: text 64, 8 x 256 tokens, S = 2115 / 2368,
batch 4), where the CPU oracle would take minutes per case.  Small-size parity with the oracle and the golden fixtures
is in test_gpu_kernels.py / test_gpu_models.py; here the same kernels are checked through invariants of the maths:
homogeneity and row-permutation equivariance of the GEMMs, convexity and key-permutation invariance of attention,
idempotence of vector quantisation, frame independence of the VQGAN decoder, determinism / decode consistency of the
mask-predict sampler, and KV-cache decode == full causal re-forward (what the reference computes, dalle_artv.py:262-290).
"""
import pytest
import torch

from cases import BERT_CASES
from helpers import build_artv, build_bert, build_vae, relerr
from mmvid_b200 import ops, synth
from mmvid_b200._lib import MASK_NONE, MASK_PREV

pytestmark = pytest.mark.gpu

B_A, S_A, D = 4, 2115, 768


def _operands(M, K, N, prec, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) * 0.02).cuda()
    if prec == "bf16":
        x, w = x.bfloat16(), w.bfloat16()
    return x, w


@pytest.mark.parametrize("prec", ["tf32", "bf16", "fp32"])
@pytest.mark.parametrize("M,K,N", [(B_A * S_A, 768, 3072), (B_A * S_A, 3072, 768), (B_A * S_A, 768, 2304)])
def test_gemm_fullsize_homogeneity_and_row_permutation_are_exact(prec, M, K, N):
    """linear(2x) == 2 linear(x) and linear(x[perm]) == linear(x)[perm], bit for bit: scaling by a power of two commutes
    with the tf32 / bf16 operand rounding and the fp32 accumulation, and every output row depends on its own input row
    only (same k-order whatever tile it lands in).  (K = 3072 runs the cta_group::2 kernel.)"""
    x, w = _operands(M, K, N, prec, 3)
    y = ops.linear(x, w, precision=prec)
    assert torch.isfinite(y).all()
    y2 = ops.linear(x * 2, w, precision=prec)
    assert torch.equal(y2, y * 2)
    perm = torch.randperm(M, generator=torch.Generator().manual_seed(1)).cuda()
    yp = ops.linear(x[perm].contiguous(), w, precision=prec)
    assert torch.equal(yp, y[perm])
    # additivity within the precision of the mode: linear(x1 + x2) ~ linear(x1) + linear(x2)
    x2, _ = _operands(M, K, N, prec, 4)
    ysum = ops.linear((x.float() + x2.float()).to(x.dtype), w, precision=prec)
    tol = {"fp32": 1e-5, "tf32": 2e-3, "bf16": 2e-2}[prec]
    assert relerr(ysum, y + ops.linear(x2, w, precision=prec)) < tol


@pytest.mark.parametrize("prec", ["tf32", "bf16"])
@pytest.mark.parametrize("kind", ["none", "prev"])
def test_attention_fullsize_rows_are_convex_combinations(prec, kind):
    """softmax rows sum to one: with V constant along the sequence the output equals that constant for EVERY query row,
    whatever the mask, the running-max rescales and the 17 key tiles; and (no mask) permuting the keys jointly with the
    values leaves the output unchanged."""
    B, S, H = B_A, S_A, 12
    g = torch.Generator().manual_seed(7)
    qkv = torch.randn(B * S, 3 * D, generator=g).cuda()
    const = torch.randn(B, 1, D, generator=g).cuda().expand(B, S, D).reshape(B * S, D)
    qkv_c = qkv.clone()
    qkv_c[:, 2 * D:] = const
    mk, rows = (MASK_NONE, []) if kind == "none" else (MASK_PREV, [65, 66])
    odt = torch.float32 if prec == "tf32" else torch.bfloat16   # the model's activation dtype of each mode
    out = ops.attention_tc(qkv_c, B, S, H, mk, rows, prec, out_dtype=odt).float()
    tol = 2e-3 if prec == "tf32" else 1e-2
    assert (out - const).abs().max() < tol * const.abs().max()
    if kind == "none":
        ref = ops.attention_tc(qkv, B, S, H, mk, rows, prec, out_dtype=odt).float()
        perm = torch.randperm(S, generator=g).cuda()
        q3 = qkv.view(B, S, 3 * D).clone()
        q3[:, :, D:] = q3[:, perm, D:]
        out_p = ops.attention_tc(q3.view(B * S, 3 * D), B, S, H, mk, rows, prec, out_dtype=odt).float()
        if prec == "bf16":   # fp32 output from the bf16 kernel (not a model path) must agree with its bf16 output
            out32 = ops.attention_tc(qkv, B, S, H, mk, rows, prec, out_dtype=torch.float32)
            assert relerr(out32, ref) < 5e-3
        assert relerr(out_p, ref) < (2e-3 if prec == "tf32" else 2e-2)


def test_vq_fullsize_quantisation_is_idempotent_and_minimal():
    """32 frames x 256 tokens against the 1024 x 256 codebook: quantising a codebook row returns that row's index
    (idempotence), and the selected code is a true nearest neighbour in float64."""
    g = torch.Generator().manual_seed(9)
    z = (torch.randn(32 * 256, 256, generator=g) * 0.3).cuda()
    cb = (torch.randn(1024, 256, generator=g) * 0.3).cuda()
    ids = ops.vq_argmin(z, cb)
    assert ids.dtype == torch.int64 and int(ids.min()) >= 0 and int(ids.max()) < 1024
    zq = ops.codebook_gather(ids, cb)
    assert torch.equal(zq, cb[ids])
    assert torch.equal(ops.vq_argmin(zq, cb), ids)
    d = torch.cdist(z.double(), cb.double()) ** 2
    picked = d.gather(1, ids[:, None])[:, 0]
    assert float((picked - d.min(1).values).max()) < 1e-4   # fp32 distance rounding may swap near-ties only


def test_vqgan_decode_fullsize_frames_are_independent():
    """Decoding a frame alone or inside a batch of 256-px frames gives the same pixels: the implicit-GEMM convolutions,
    GroupNorm statistics and the mid attention are per-image (model.py:551-582)."""
    for prec, tol in (("fp32", 1e-6), ("tf32", 1e-6)):
        vae, _ = build_vae(256, 77, precision=prec)
        ids = synth.synth_codes(4, 256, 1024, 5).cuda()
        full = vae.decode(ids)
        one = vae.decode(ids[2:3].contiguous())
        assert full.shape == (4, 3, 256, 256) and float(full.min()) >= 0.0 and float(full.max()) <= 1.0
        assert (full[2:3] - one).abs().max() <= tol, prec


def test_bert_shapeA_generation_is_seeded_valid_and_decode_consistent():
    """BERT.generate_images at Shape A, batch 4 (the bench workload): same seed -> same ids; ids in the image vocabulary;
    frames in [0, 1]; the returned frames are exactly the VQGAN decode of the returned ids (dalle_bert.py:476-487)."""
    cfg = BERT_CASES["bert_shapeA"]
    model, _ = build_bert(cfg, precision="tf32", sampling_mode="batched")
    text = synth.synth_text(B_A, cfg["text_seq_len"], cfg["vocab"], 3).cuda()
    outs = []
    for _ in range(2):
        torch.manual_seed(123)
        images, _, seq = model.generate_images(text, mask_predict_steps=4, dynamic=False)
        outs.append((images, seq))
    (im0, s0), (im1, s1) = outs
    assert torch.equal(s0, s1) and torch.equal(im0, im1)
    assert s0.shape == (B_A * 8, 256) and int(s0.min()) >= 0 and int(s0.max()) < 1024
    assert im0.shape == (B_A, 8, 3, 256, 256) and float(im0.min()) >= 0.0 and float(im0.max()) <= 1.0
    again = model.vae.decode(s0).view_as(im0)
    assert (again - im0).abs().max() <= 1e-6
    torch.manual_seed(124)
    _, _, s2 = model.generate_images(text, mask_predict_steps=4, dynamic=False)
    assert not torch.equal(s2, s0)


@pytest.mark.parametrize("impl", ["fused", "native", "persistent"])
def test_artv_shapeA_kv_cache_logits_match_full_causal_forward(impl):
    """ART-V at Shape A (prefix 321, 2048 decode steps, S = 2368): the per-step image logits produced from the KV cache
    equal the rows of ONE full causal forward over the generated sequence - which is what the reference recomputes from
    scratch for every token (dalle_artv.py:262-290)."""
    cfg = dict(dim=768, layers=12, text_seq_len=64, vocab=49408, num_visuals=1, num_targets=8, image_size=256, seed=35,
               batch=2)
    model, _ = build_artv(cfg, precision="fp32", sampling_mode="batched")
    model.decode_impl = impl
    B = cfg["batch"]
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], 4).cuda()
    visual = synth.synth_frames(B, 1, cfg["image_size"], 6).cuda()
    trace = []
    torch.manual_seed(5)
    _, _, toks = model.generate_images(text, visual=visual, return_tokens=True, logits_trace=trace)
    assert toks.shape == (B, 2048) and len(trace) == 2048
    step_logits = torch.stack(trace, 1)                         # [B, 2048, 1024]
    P = cfg["text_seq_len"] + 1 + 256
    lo = model.num_control_tokens
    full = model(text, visual=visual, target=toks)              # [B, 2368, total_tokens]
    rows = full[:, P - 1:P - 1 + 2048, lo:lo + 1024]
    assert rows.shape == step_logits.shape
    assert relerr(step_logits, rows) < 1e-4
    assert float(full[:, P - 1:, :lo].max()) < -1e30            # text / visual vocabulary masked in image rows


@pytest.mark.parametrize("prec,tol", [("fp16", 2e-3), ("tf32", 2e-3), ("bf16", 3e-2)])
def test_artv_shapeA_streaming_decode_logits_match_full_causal_forward(prec, tol):
    """The persistent streaming decode kernel (decode_stream.cu: one launch per token, 16-bit weights and K/V cache, fp32
    accumulation) at Shape A, batch 4 (BASELINE config 3): per-step image logits from the cache against ONE full causal
    forward of the same model in the same precision over the generated sequence, and against the fp32 path's forward."""
    cfg = dict(dim=768, layers=12, text_seq_len=64, vocab=49408, num_visuals=1, num_targets=8, image_size=256, seed=35,
               batch=4)
    model, _ = build_artv(cfg, precision=prec, sampling_mode="batched")
    assert getattr(model, "decode_impl", None) is None  # default = streaming kernel in the tensor-core precisions
    B = cfg["batch"]
    text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], 4).cuda()
    visual = synth.synth_frames(B, 1, cfg["image_size"], 6).cuda()
    trace = []
    torch.manual_seed(5)
    _, _, toks = model.generate_images(text, visual=visual, return_tokens=True, logits_trace=trace)
    assert toks.shape == (B, 2048) and len(trace) == 2048 and int(toks.min()) >= 0 and int(toks.max()) < 1024
    step_logits = torch.stack(trace, 1)
    P = cfg["text_seq_len"] + 1 + 256
    lo = model.num_control_tokens
    full = model(text, visual=visual, target=toks)
    rows = full[:, P - 1:P - 1 + 2048, lo:lo + 1024]
    e_same = relerr(step_logits, rows)
    model.precision = model.transformer.precision = "fp32"
    rows32 = model(text[:1], visual=visual[:1], target=toks[:1])[:, P - 1:P - 1 + 2048, lo:lo + 1024]
    e_fp32 = relerr(step_logits[:1], rows32)
    print(f"artv shape A streaming decode {prec}: logits relerr vs own full forward {e_same:.2e}, vs fp32 forward {e_fp32:.2e}")
    assert e_same < tol and e_fp32 < tol


_PDL_PROBE = r"""
import hashlib, sys, torch
sys.path[:0] = [{root!r}, {root!r} + "/tests", {root!r} + "/tests/golden"]
from cases import BERT_CASES
from helpers import build_bert
from mmvid_b200 import synth
cfg = BERT_CASES["bert_shapeB"]
model, _ = build_bert(cfg, precision="fp16")
B = cfg["batch"]
text = synth.synth_text(B, cfg["text_seq_len"], cfg["vocab"], cfg["seed"]).cuda()
torch.manual_seed(5)
images, _, seq = model.generate_images(text, visual=None, mask_predict_steps=3, dynamic=False)
h = hashlib.sha256(seq.cpu().numpy().tobytes() + images.cpu().numpy().tobytes()).hexdigest()
print("PDLHASH", h)
"""


def test_programmatic_dependent_launch_does_not_change_a_single_bit():
    """MMVID_PDL=1 (default) lets LayerNorm / attention / GEMM kernels start while their predecessor drains; every one of them
    must wait (griddepcontrol.wait) before it touches global memory.  A missing wait is a race, and a race shows up as a
    different bit somewhere: sampled ids and decoded frames of a full generate_images call (768 x 12 transformer, CUDA
    graph, VQGAN decode) must hash identically with the chaining on and off.  (The switch is read once per process.)"""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hashes = {}
    for pdl in ("0", "1", "1"):
        env = dict(os.environ, MMVID_PDL=pdl)
        r = subprocess.run([sys.executable, "-c", _PDL_PROBE.format(root=root)], env=env, capture_output=True, text=True, timeout=600)
        lines = [ln for ln in r.stdout.splitlines() if ln.startswith("PDLHASH ")]
        assert r.returncode == 0 and lines, r.stderr[-2000:]
        hashes.setdefault(pdl, []).append(lines[-1].split()[1])
    print("PDL off / on / on:", hashes)
    assert len(set(hashes["0"] + hashes["1"])) == 1


def test_pinned_stager_round_trip_and_double_buffering():
    """tokenizer.PinnedStager: batches come back in order and intact, slots are reused after two batches."""
    from mmvid_b200.tokenizer import PinnedStager
    st = PinnedStager("cuda", slots=2)
    g = torch.Generator().manual_seed(1)
    batches = [dict(text=torch.randint(0, 1000, (4, 64), generator=g), frames=torch.rand(4, 2, 3, 32, 32, generator=g), visuals=None)
               for _ in range(5)]
    st.put(**batches[0])
    for i in range(5):
        if i + 1 < 5:
            st.put(**batches[i + 1])  # next batch's copy is in flight while this one is consumed
        got = st.get()
        assert got["visuals"] is None and got["text"].is_cuda
        assert torch.equal(got["text"].cpu(), batches[i]["text"]) and torch.equal(got["frames"].cpu(), batches[i]["frames"])
