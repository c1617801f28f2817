"""Torch-tensor front end of the C ABI: argument checking, scratch allocation, stream plumbing.

PyTorch is used for device memory, streams and RNG only; every FLOP of the hot path runs in
libmmvid_b200.so.  All functions require CUDA tensors and raise if the library is missing.
"""
import ctypes as C

import torch

from . import _lib as L
from ._lib import (ACT_NONE, ACT_QUICKGELU, ACT_SWISH, BF16, DT_BF16, DT_F16, DT_F32, F16, FP32, H16, MASK_CAUSAL,  # noqa: F401
                   MASK_NONE, MASK_PREV, TF32)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _req(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"mmvid_b200: {name} must be a CUDA tensor (no CPU fallback exists)")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"mmvid_b200: {name} must be {dtype}, got {t.dtype}")
    return t


def _dt(t):
    if t.dtype == torch.float32:
        return DT_F32
    if t.dtype == torch.bfloat16:
        return DT_BF16
    if t.dtype == torch.float16:
        return DT_F16
    raise RuntimeError(f"unsupported dtype {t.dtype}")


def act_dtype(precision):
    """torch dtype of GEMM operands / inter-kernel activations for a precision (fp32 for FP32 and TF32)."""
    precision = precision_id(precision)
    return torch.bfloat16 if precision == BF16 else (torch.float16 if precision == F16 else torch.float32)


def precision_id(p):
    if isinstance(p, int):
        return p
    return L.PRECISIONS[p]


# ------------------------------------------------------------------------------------------------ embedding

def embed_gather(out, segments):
    """out: float32 [B,S,D]; segments: list of dict(ids=int64[B,n] (may be a broadcast view with stride 0),
    seq_off, table, table2=None, pos=None, pad=None|(pad_value, pad_base))."""
    lib = L.load()
    _req(out, torch.float32, "out")
    B, S, D = out.shape
    arr = (L.EmbedSegment * len(segments))()
    keep = []
    for i, s in enumerate(segments):
        ids = _req(s["ids"], torch.int64, "ids")
        assert ids.dim() == 2 and ids.stride(1) == 1 and ids.shape[0] in (1, B)
        bstride = ids.stride(0) if ids.shape[0] == B else 0
        keep.append(ids)
        e = arr[i]
        e.ids, e.ids_bstride, e.n, e.seq_off = ids.data_ptr(), bstride, ids.shape[1], s["seq_off"]
        e.table = _req(s["table"], torch.float32, "table").data_ptr()
        assert s["table"].is_contiguous() and s["table"].shape[1] == D
        e.table_rows = s["table"].shape[0]
        assert s.get("table2") is None or s["table2"].shape[0] >= e.table_rows
        e.table2 = s["table2"].data_ptr() if s.get("table2") is not None else None
        e.pos = s["pos"].data_ptr() if s.get("pos") is not None else None
        if s.get("pos") is not None:
            assert s["pos"].is_contiguous() and s["pos"].shape[0] >= ids.shape[1]
        if s.get("pad") is not None:
            e.pad_value, e.pad_base, e.use_pad = s["pad"][0], s["pad"][1], 1
        else:
            e.pad_value, e.pad_base, e.use_pad = 0, 0, 0
    assert out.is_contiguous()
    L.check(lib.mmvid_embed_gather(_ptr(out), B, S, D, arr, len(segments), _stream()), "embed_gather")
    return out


def axial_table(weights, shape, n=None):
    """Sum of broadcast axial tables -> float32 [n, D].  weights: list of tensors shaped (1, .., n_i, .., 1, D)."""
    lib = L.load()
    D = weights[0].shape[-1]
    total = 1
    for s in shape:
        total *= s
    n = total if n is None else n
    ws = [_req(w, torch.float32).reshape(-1, D).contiguous() for w in weights]
    out = torch.empty(n, D, device=ws[0].device, dtype=torch.float32)
    sh = (C.c_int * 3)(*(list(shape) + [1] * (3 - len(shape))))
    w = ws + [None] * (3 - len(ws))
    L.check(lib.mmvid_axial_table(_ptr(out), n, D, _ptr(w[0]), _ptr(w[1]), _ptr(w[2]), sh, len(shape), _stream()),
            "axial_table")
    return out


# ------------------------------------------------------------------------------------------------ norms

def layernorm(x, weight, bias, eps=1e-5, out_dtype=torch.float32):
    lib = L.load()
    _req(x, torch.float32, "x")
    D = x.shape[-1]
    assert x.stride(-1) == 1
    x2 = x.reshape(-1, D)
    assert x2.stride(1) == 1
    out = torch.empty(x2.shape, device=x.device, dtype=out_dtype)
    L.check(lib.mmvid_layernorm(_ptr(x2), x2.stride(0), _ptr(weight), _ptr(bias), _ptr(out), _dt(out), x2.shape[0], D,
                                eps, _stream()), "layernorm")
    return out.view(*x.shape[:-1], D)


def groupnorm(x_nhwc, weight, bias, groups=32, eps=1e-6, swish=False, out=None, fast=False, out_dtype=torch.float32,
              partial=None):
    """x: float32 [N, H, W, C] contiguous.  fast: swish through MUFU ex2 / rcp (tensor-core precision modes).
    out_dtype=torch.float16: the result feeds a kind::f16 conv (not in place).
    partial: per-slab (sum, squared deviations) the conv that produced x wrote next to it (conv2d(..., gn_groups=)): the
    statistics pass over x is skipped."""
    lib = L.load()
    _req(x_nhwc, torch.float32, "x")
    assert x_nhwc.is_contiguous()
    N, H, W, Cc = x_nhwc.shape
    if out is None or out.dtype != out_dtype:
        out = torch.empty(x_nhwc.shape, device=x_nhwc.device, dtype=out_dtype)
    if partial is not None:
        assert partial.dtype == torch.float32 and partial.numel() == N * H * W // 32 * groups * 2
        stats = torch.empty(N * groups * 2, device=x_nhwc.device, dtype=torch.float32)
        L.check(lib.mmvid_groupnorm_from_partials(_ptr(x_nhwc), _ptr(out), _dt(out), _ptr(weight), _ptr(bias), _ptr(partial),
                                                  _ptr(stats), N, H * W, Cc, groups, eps, (2 if fast else 1) if swish else 0,
                                                  _stream()), "groupnorm_from_partials")
        return out
    stats = torch.empty(int(lib.mmvid_groupnorm_scratch_floats(N, groups)), device=x_nhwc.device, dtype=torch.float32)
    L.check(lib.mmvid_groupnorm(_ptr(x_nhwc), _ptr(out), _dt(out), _ptr(weight), _ptr(bias), _ptr(stats), N, H * W, Cc,
                                groups, eps, (2 if fast else 1) if swish else 0, _stream()), "groupnorm")
    return out


# ------------------------------------------------------------------------------------------------ linear / gemm

def linear(a, w, bias=None, act=ACT_NONE, residual=None, precision=FP32, out=None, out_dtype=None):
    """C[M,N] = act(A[M,K] W[N,K]^T + bias) (+ residual).  a, w: fp32 (FP32/TF32) or bf16 (BF16)."""
    lib = L.load()
    precision = precision_id(precision)
    K = a.shape[-1]
    a2 = a.reshape(-1, K)
    assert a2.stride(1) == 1 and w.stride(1) == 1 and w.shape[1] == K
    M, N = a2.shape[0], w.shape[0]
    if out_dtype is None:
        out_dtype = torch.float32
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    o2 = out.reshape(M, N)
    assert o2.stride(1) == 1
    r2 = residual.reshape(M, N) if residual is not None else None
    L.check(lib.mmvid_linear(_ptr(a2), _dt(a2), a2.stride(0), _ptr(w), _dt(w), w.stride(0), _ptr(bias), _ptr(r2),
                             r2.stride(0) if r2 is not None else 0, _ptr(o2), _dt(o2), o2.stride(0), M, N, K, act,
                             precision, _stream()), "linear")
    return out.view(*a.shape[:-1], N) if out.dim() == 2 and a.dim() != 2 else out


def linear_small_m(a, w, bias=None, act=ACT_NONE, residual=None, out=None):
    lib = L.load()
    M, K = a.shape
    N = w.shape[0]
    if M > 16:
        return linear(a, w, bias, act=act, residual=residual, precision=FP32, out=out)
    out = torch.empty(M, N, device=a.device, dtype=torch.float32) if out is None else out
    L.check(lib.mmvid_linear_small_m(_ptr(a), a.stride(0), _ptr(w), w.stride(0), _ptr(bias), _ptr(residual),
                                     residual.stride(0) if residual is not None else 0, _ptr(out), out.stride(0), M, N,
                                     K, act, _stream()), "linear_small_m")
    return out


def gemm_batched_f32(A, lda, a_s1, a_s2, Bm, ldb_n, ldb_k, b_s1, b_s2, Cm, ldc, c_s1, c_s2, M, N, K, batch1, batch2,
                     alpha=1.0):
    lib = L.load()
    L.check(lib.mmvid_gemm_batched_f32(_ptr(A), lda, a_s1, a_s2, _ptr(Bm), ldb_n, ldb_k, b_s1, b_s2, _ptr(Cm), ldc,
                                       c_s1, c_s2, M, N, K, batch1, batch2, alpha, _stream()), "gemm_batched_f32")


def softmax_rows(scores, rows, cols, ld, batch, mask_kind=MASK_NONE, prev_rows=None):
    lib = L.load()
    n_prev = 0 if prev_rows is None else prev_rows.numel()
    L.check(lib.mmvid_softmax_rows(_ptr(scores), batch, rows, cols, ld, mask_kind, _ptr(prev_rows), n_prev, _stream()),
            "softmax_rows")


def softmax_logits(logits, noise=None, noise_scale=0.0):
    lib = L.load()
    _req(logits, torch.float32)
    n = logits.shape[-1]
    l2 = logits.reshape(-1, n)
    assert l2.is_contiguous()
    probs = torch.empty_like(l2)
    if noise is not None:
        assert noise.is_contiguous() and noise.numel() == l2.numel()
    L.check(lib.mmvid_softmax_logits(_ptr(l2), _ptr(noise), float(noise_scale), _ptr(probs), l2.shape[0], n, _stream()),
            "softmax_logits")
    return probs.view_as(logits)


def mp_sample(logits, Y, tok, seed, offset, noise_scale=0.0, skip=None, step_dev=None):
    """Fused softmax + categorical draw + gather per row of `logits` [..., n] (dalle_bert.py:527-534); updates Y (float32)
    and tok (int64) of the rows whose `skip` flag (uint8 / bool) is not set."""
    lib = L.load()
    _req(logits, torch.float32)
    n = logits.shape[-1]
    l2 = logits.reshape(-1, n)
    assert l2.is_contiguous() and Y.is_contiguous() and tok.is_contiguous() and Y.numel() == l2.shape[0] == tok.numel()
    assert Y.dtype == torch.float32 and tok.dtype == torch.int64
    if skip is not None:
        skip = skip.view(torch.uint8) if skip.dtype == torch.bool else skip
        assert skip.is_contiguous() and skip.numel() == l2.shape[0]
    L.check(lib.mmvid_mp_sample(_ptr(l2), l2.shape[0], n, float(noise_scale), _ptr(skip), _ptr(Y), _ptr(tok), int(seed),
                                int(offset), _ptr(step_dev), _stream()), "mp_sample")


def mp_keep(Y, pmask, I_tok, k, mask_id, seed, offset, beams=1):
    """Keep-mask draw of one mask-predict iteration (dalle_bert.py:646-664): keep k not-preserved tokens per (sample, beam)
    without replacement, proportional to Y.  Returns (keep bool [samples*beams, Ttot], ids_in int64 [samples*beams, Ttot])."""
    lib = L.load()
    _req(Y, torch.float32)
    nb, Ttot = Y.shape
    assert Y.is_contiguous() and I_tok.is_contiguous() and I_tok.dtype == torch.int64 and I_tok.shape == Y.shape
    pm = None
    if pmask is not None:
        pm = pmask.reshape(-1)
        pm = pm.view(torch.uint8) if pm.dtype == torch.bool else pm
        assert pm.numel() == Ttot and pm.is_contiguous()
    keep = torch.empty(nb * beams, Ttot, dtype=torch.uint8, device=Y.device)
    ids_in = torch.empty(nb * beams, Ttot, dtype=torch.int64, device=Y.device)
    L.check(lib.mmvid_mp_keep(_ptr(Y), _ptr(pm), _ptr(I_tok), nb, beams, Ttot, int(k), int(mask_id), _ptr(keep), _ptr(ids_in),
                              int(seed), int(offset), _stream()), "mp_keep")
    return keep.view(torch.bool), ids_in


# ------------------------------------------------------------------------------------------------ attention

def attention_fp32(qkv, B, S, H, mask_kind, prev_rows_dev):
    """CUDA-core parity path: qkv float32 [B*S, 3*H*64] -> out float32 [B*S, H*64].
    QK^T and PV as strided-batched fp32 GEMMs around an in-place masked row softmax."""
    D = H * 64
    out = torch.empty(B * S, D, device=qkv.device, dtype=torch.float32)
    # one batch element at a time keeps the [H,S,S] score scratch bounded (215 MB at S=2115)
    scores = torch.empty(H, S, S, device=qkv.device, dtype=torch.float32)
    for b in range(B):
        base = qkv[b * S:(b + 1) * S]
        q, k, v = base[:, :D], base[:, D:2 * D], base[:, 2 * D:]
        gemm_batched_f32(q, 3 * D, 0, 64, k, 3 * D, 1, 0, 64, scores, S, 0, S * S, S, S, 64, 1, H, alpha=0.125)
        softmax_rows(scores, S, S, S, H, mask_kind, prev_rows_dev)
        ob = out[b * S:(b + 1) * S]
        gemm_batched_f32(scores, S, 0, S * S, v, 1, 3 * D, 0, 64, ob, D, 0, 64, S, 64, S, 1, H)
    return out


def attention_tc(qkv, B, S, H, mask_kind, prev_rows_host, precision, out_dtype=torch.float32):
    """Tensor-core path: split qkv into Q,K [B,H,S_pad,64] and V^T [B,H,64,S_pad], then fused flash attention."""
    lib = L.load()
    precision = precision_id(precision)
    S_pad = (S + 127) // 128 * 128
    dt = act_dtype(precision)
    dev = qkv.device
    q = torch.empty(B, H, S_pad, 64, device=dev, dtype=dt)
    k = torch.empty(B, H, S_pad, 64, device=dev, dtype=dt)
    vt = torch.empty(B, H, 64, S_pad, device=dev, dtype=dt)
    L.check(lib.mmvid_qkv_split(_ptr(qkv), _ptr(q), _ptr(k), _ptr(vt), _dt(q), B, H, S, S_pad, _stream()), "qkv_split")
    out = torch.empty(B * S, H * 64, device=dev, dtype=out_dtype)
    pr = (C.c_int * 4)(*(list(prev_rows_host) + [0] * (4 - len(prev_rows_host))))
    L.check(lib.mmvid_attention(_ptr(q), _ptr(k), _ptr(vt), _ptr(out), _dt(out), out.stride(0), B, H, S, S_pad,
                                mask_kind, pr, len(prev_rows_host), precision, _stream()), "attention")
    return out


def alloc_qkv_buffers(B, H, S, precision, device):
    """Zero-initialised Q,K [B,H,S_pad,64] and V^T [B,H,64,S_pad] (padding must stay zero: P x garbage = NaN)."""
    precision = precision_id(precision)
    S_pad = (S + 127) // 128 * 128
    dt = act_dtype(precision)
    q = torch.zeros(B, H, S_pad, 64, device=device, dtype=dt)
    k = torch.zeros(B, H, S_pad, 64, device=device, dtype=dt)
    vt = torch.zeros(B, H, 64, S_pad, device=device, dtype=dt)
    return q, k, vt


def linear_qkv(a, w, bias, bufs, B, S, H, precision):
    """Fused in-projection: writes Q,K,V^T of `a @ w.T + bias` straight into the attention layout."""
    lib = L.load()
    precision = precision_id(precision)
    q, k, vt = bufs
    S_pad = q.shape[2]
    a2 = a.reshape(B * S, H * 64)
    assert a2.stride(1) == 1 and w.stride(1) == 1 and w.shape == (3 * H * 64, H * 64)
    L.check(lib.mmvid_linear_qkv(_ptr(a2), _dt(a2), a2.stride(0), _ptr(w), _dt(w), w.stride(0), _ptr(bias), _ptr(q), _ptr(k),
                                 _ptr(vt), _dt(q), B, H, S, S_pad, precision, _stream()), "linear_qkv")


def attention_core(bufs, B, S, H, mask_kind, prev_rows_host, precision, out_dtype=torch.float32):
    lib = L.load()
    precision = precision_id(precision)
    q, k, vt = bufs
    S_pad = q.shape[2]
    out = torch.empty(B * S, H * 64, device=q.device, dtype=out_dtype)
    pr = (C.c_int * 4)(*(list(prev_rows_host) + [0] * (4 - len(prev_rows_host))))
    L.check(lib.mmvid_attention(_ptr(q), _ptr(k), _ptr(vt), _ptr(out), _dt(out), out.stride(0), B, H, S, S_pad,
                                mask_kind, pr, len(prev_rows_host), precision, _stream()), "attention")
    return out


def kv_append(qkv, kcache, vcache, pos):
    lib = L.load()
    B, H, S_max, _ = kcache.shape
    L.check(lib.mmvid_kv_append(_ptr(qkv), qkv.stride(0), _ptr(kcache), _ptr(vcache), B, H, S_max, pos, _stream()),
            "kv_append")


def decode_attention(qkv, kcache, vcache, length):
    lib = L.load()
    B, H, S_max, _ = kcache.shape
    out = torch.empty(B, H * 64, device=qkv.device, dtype=torch.float32)
    L.check(lib.mmvid_decode_attention(_ptr(qkv), qkv.stride(0), _ptr(kcache), _ptr(vcache), _ptr(out), out.stride(0),
                                       B, H, S_max, length, _stream()), "decode_attention")
    return out


# ------------------------------------------------------------------------------------------------ VQ / conv

def vq_argmin(z_rows, codebook):
    lib = L.load()
    _req(z_rows, torch.float32)
    assert z_rows.is_contiguous() and codebook.is_contiguous()
    T, dim = z_rows.shape
    idx = torch.empty(T, device=z_rows.device, dtype=torch.int64)
    e2 = torch.empty(codebook.shape[0], device=z_rows.device, dtype=torch.float32)  # stream-ordered scratch
    L.check(lib.mmvid_vq_argmin(_ptr(z_rows), _ptr(codebook), _ptr(idx), _ptr(e2), T, codebook.shape[0], dim,
                                _stream()), "vq_argmin")
    return idx


def codebook_gather(ids, codebook):
    lib = L.load()
    _req(ids, torch.int64)
    ids = ids.contiguous().view(-1)
    out = torch.empty(ids.numel(), codebook.shape[1], device=ids.device, dtype=torch.float32)
    L.check(lib.mmvid_codebook_gather(_ptr(ids), _ptr(codebook), _ptr(out), ids.numel(), codebook.shape[1], _stream()),
            "codebook_gather")
    return out


def conv2d(x, w_packed, bias, *, stride=1, pad=(1, 1), out_hw=None, upsample=False, residual=None, in_nchw=False,
           out_nchw=False, pre_affine=False, post_clamp=False, precision=FP32, gn_groups=0):
    """x: float32 NHWC [N,H,W,Cin] (or NCHW if in_nchw); w_packed [Cout,KH,KW,Cin]; returns float32 NHWC (or NCHW).
    precision 'fp16': x and w_packed are float16 (kind::f16 implicit GEMM, fp32 accumulate / bias / residual / output).
    gn_groups > 0: returns (out, partial) where partial holds the GroupNorm partial statistics of `out` for
    groupnorm(..., partial=) - or None where the conv cannot produce them (mmvid_conv2d_gn_fusable)."""
    lib = L.load()
    if precision_id(precision) == F16:
        _req(x, torch.float16)
        _req(w_packed, torch.float16, "w_packed")
    else:
        _req(x, torch.float32)
    assert x.is_contiguous() and w_packed.is_contiguous()
    if in_nchw:
        N, Cin, H, W = x.shape
    else:
        N, H, W, Cin = x.shape
    Cout, KH, KW, Cin2 = w_packed.shape
    assert Cin2 == Cin
    Hs, Ws = (2 * H, 2 * W) if upsample else (H, W)
    if out_hw is None:
        Ho = (Hs + 2 * pad[0] - KH) // stride + 1
        Wo = (Ws + 2 * pad[1] - KW) // stride + 1
    else:
        Ho, Wo = out_hw
    out = torch.empty((N, Cout, Ho, Wo) if out_nchw else (N, Ho, Wo, Cout), device=x.device, dtype=torch.float32)
    p = L.ConvParams()
    p.inp, p.w, p.bias, p.out = x.data_ptr(), w_packed.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr()
    p.residual = residual.data_ptr() if residual is not None else None
    if residual is not None:
        assert residual.is_contiguous() and residual.shape == out.shape
    p.N, p.H, p.W, p.Cin, p.Cout, p.KH, p.KW = N, H, W, Cin, Cout, KH, KW
    p.stride, p.pad_t, p.pad_l, p.Ho, p.Wo = stride, pad[0], pad[1], Ho, Wo
    p.upsample, p.in_nchw, p.out_nchw = int(upsample), int(in_nchw), int(out_nchw)
    p.pre_affine, p.post_clamp, p.precision = int(pre_affine), int(post_clamp), precision_id(precision)
    partial = None
    if gn_groups:
        p.gn_groups = int(gn_groups)
        if lib.mmvid_conv2d_gn_fusable(C.byref(p)) == 1:
            partial = torch.empty(N * Ho * Wo // 32 * gn_groups * 2, device=x.device, dtype=torch.float32)
            p.gn_partial = partial.data_ptr()
    L.check(lib.mmvid_conv2d(C.byref(p), _stream()), "conv2d")
    return (out, partial) if gn_groups else out


def upsample2x(x, out_dtype=torch.float32):
    """nearest x2 on NHWC float32 (model.py:57-59); out_dtype=torch.float16 when a kind::f16 conv consumes it."""
    lib = L.load()
    _req(x, torch.float32)
    N, H, W, Cc = x.shape
    out = torch.empty(N, 2 * H, 2 * W, Cc, device=x.device, dtype=out_dtype)
    L.check(lib.mmvid_upsample2x(_ptr(x), _ptr(out), _dt(out), N, H, W, Cc, _stream()), "upsample2x")
    return out


def conv_out_fused(x_nhwc, gamma, beta, w_packed, bias, groups=32, eps=1e-6, post_clamp=True, fast=False, partial=None):
    """GroupNorm + swish + 3x3 conv (Cout <= 4) + clamp/rescale; NHWC float32 in, NCHW out.
    partial: GroupNorm partial statistics of x written by the conv that produced it (conv2d(..., gn_groups=))."""
    lib = L.load()
    _req(x_nhwc, torch.float32)
    assert x_nhwc.is_contiguous() and w_packed.is_contiguous()
    N, H, W, Cc = x_nhwc.shape
    Cout = w_packed.shape[0]
    assert tuple(w_packed.shape[1:]) == (3, 3, Cc)
    out = torch.empty(N, Cout, H, W, device=x_nhwc.device, dtype=torch.float32)
    if partial is not None:
        assert partial.dtype == torch.float32 and partial.numel() == N * H * W // 32 * groups * 2
        stats = torch.empty(N * groups * 2, device=x_nhwc.device, dtype=torch.float32)
        L.check(lib.mmvid_conv_out_fused_from_partials(_ptr(x_nhwc), _ptr(gamma), _ptr(beta), _ptr(w_packed), _ptr(bias), _ptr(out),
                                                       _ptr(partial), _ptr(stats), N, H, W, Cc, Cout, groups, eps,
                                                       int(post_clamp) | (2 if fast else 0), _stream()), "conv_out_fused_from_partials")
        return out
    stats = torch.empty(int(lib.mmvid_groupnorm_scratch_floats(N, groups)), device=x_nhwc.device, dtype=torch.float32)
    L.check(lib.mmvid_conv_out_fused(_ptr(x_nhwc), _ptr(gamma), _ptr(beta), _ptr(w_packed), _ptr(bias), _ptr(out),
                                     _ptr(stats), N, H, W, Cc, Cout, groups, eps, int(post_clamp) | (2 if fast else 0), _stream()),
            "conv_out_fused")
    return out
