"""Training path: torch.autograd.Function wrappers whose forward AND backward run in libmmvid_b200.so.

Used by `BERT.forward(return_loss=True)` (dalle_bert.py:980-1127) so that `loss.backward()`, `clip_grad_norm_`,
Adam and DDP from the reference's train.py:320-325 work on the drop-in module unchanged.  Forward math is the same
kernels as inference (tensor-core GEMMs, flash attention); the backward is assembled from
  * tensor-core GEMMs on explicitly transposed operands (dX = dY W, dW = dY^T X),
  * a recompute-based fp32 attention backward (scores/softmax rebuilt per batch element from the saved Q,K,V),
  * small HBM-bound kernels (LayerNorm / QuickGELU / softmax / cross-entropy backward, bias column sums,
    embedding scatter-add).
This is the first, correctness-first training path (SURVEY.md §8f rank 1 lists the fused flash-attention backward
and fused optimizer as the follow-up).
"""
import ctypes as C

import torch

from . import _lib as L
from . import ops
from ._lib import ACT_NONE, FP32, MASK_NONE, TF32


def _s():
    return ops._stream()


def _p(t):
    return ops._ptr(t)


def _prec_for(prec, k):
    """Tensor-core GEMMs need 16-byte aligned rows (K % 4 == 0 for tf32); tiny/odd shapes use the fp32 path."""
    return prec if (prec == FP32 or k % 4 == 0) else FP32


def transpose2d(x):
    lib = L.load()
    R, Cn = x.shape
    assert x.is_contiguous()
    out = torch.empty(Cn, R, device=x.device, dtype=torch.float32)
    L.check(lib.mmvid_transpose2d(_p(x), _p(out), R, Cn, _s()), "transpose2d")
    return out


def colsum(x, out=None, accumulate=False):
    lib = L.load()
    rows, cols = x.shape
    out = torch.empty(cols, device=x.device, dtype=torch.float32) if out is None else out
    scratch = torch.empty(64 * cols, device=x.device, dtype=torch.float32)
    L.check(lib.mmvid_colsum(_p(x), _p(out), _p(scratch), rows, cols, int(accumulate), _s()), "colsum")
    return out


class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b); x [M,K] fp32, W [N,K]."""

    @staticmethod
    def forward(ctx, x, w, b, act, precision):
        lib = L.load()
        x = x.contiguous()
        z = ops.linear(x, w.detach(), b.detach() if b is not None else None, precision=_prec_for(precision, x.shape[1]))
        if act != ACT_NONE:
            y = torch.empty_like(z)
            L.check(lib.mmvid_act_forward(_p(z), _p(y), z.numel(), act, _s()), "act_forward")
        else:
            y = z
        ctx.save_for_backward(x, w, z if act != ACT_NONE else None)
        ctx.act, ctx.precision, ctx.has_bias = act, precision, b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        x, w, z = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.act != ACT_NONE:
            dz = torch.empty_like(dy)
            L.check(lib.mmvid_act_backward(_p(z), _p(dy), _p(dz), dy.numel(), ctx.act, _s()), "act_backward")
        else:
            dz = dy
        prec = ctx.precision
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.linear(dz, transpose2d(w.detach()), precision=_prec_for(prec, dz.shape[1]))   # [M,N] x [N,K]
        if ctx.needs_input_grad[1]:
            dw = ops.linear(transpose2d(dz), transpose2d(x), precision=_prec_for(prec, dz.shape[0]))  # [N,M] x [M,K]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dz)
        return dx, dw, db, None, None


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous()
        y = ops.layernorm(x, gamma.detach(), beta.detach(), eps)
        ctx.save_for_backward(x, gamma)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = L.load()
        x, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        rows, D = x.shape
        dx = torch.empty_like(x)
        xhat_dy = torch.empty_like(x)
        L.check(lib.mmvid_layernorm_backward(_p(x), _p(gamma.detach()), _p(dy), _p(dx), _p(xhat_dy), rows, D, ctx.eps, _s()),
                "layernorm_backward")
        return dx, colsum(xhat_dy), colsum(dy), None


class AttentionFn(torch.autograd.Function):
    """Multi-head attention core on qkv [B*S, 3*H*64] -> [B*S, H*64].  Forward: tensor-core flash kernel
    (or the fp32 path); backward: recompute P per batch element and run the five fp32 batched GEMMs."""

    @staticmethod
    def forward(ctx, qkv, B, S, H, mask_kind, mask_rows, precision):
        qkv = qkv.contiguous()
        if precision == FP32:
            rows = torch.tensor(list(mask_rows) or [0], dtype=torch.int32, device=qkv.device)
            out = ops.attention_fp32(qkv, B, S, H, mask_kind, rows)
        else:
            out = ops.attention_tc(qkv, B, S, H, mask_kind, mask_rows, precision)
        ctx.save_for_backward(qkv)
        ctx.cfg = (B, S, H, mask_kind, tuple(mask_rows))
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = L.load()
        (qkv,) = ctx.saved_tensors
        B, S, H, mask_kind, mask_rows = ctx.cfg
        D = H * 64
        dev = qkv.device
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        rows = torch.tensor(list(mask_rows) or [0], dtype=torch.int32, device=dev)
        P = torch.empty(H, S, S, device=dev, dtype=torch.float32)
        dP = torch.empty(H, S, S, device=dev, dtype=torch.float32)
        g = ops.gemm_batched_f32
        for b in range(B):
            base, dbase, do = qkv[b * S:(b + 1) * S], dqkv[b * S:(b + 1) * S], dout[b * S:(b + 1) * S]
            q, k, v = base[:, :D], base[:, D:2 * D], base[:, 2 * D:]
            dq, dk, dv = dbase[:, :D], dbase[:, D:2 * D], dbase[:, 2 * D:]
            # P = softmax(q k^T / 8 + mask)
            g(q, 3 * D, 0, 64, k, 3 * D, 1, 0, 64, P, S, 0, S * S, S, S, 64, 1, H, alpha=0.125)
            ops.softmax_rows(P, S, S, S, H, mask_kind, rows)
            # dP = dO V^T          A = dO [S,64] (ld D), B(k=d, n=key) = v[key, d]
            g(do, D, 0, 64, v, 3 * D, 1, 0, 64, dP, S, 0, S * S, S, S, 64, 1, H)
            # dV = P^T dO          computed as dV[key, d] = sum_q P[q,key] dO[q,d]:  A(m=key,k=q) = P[q,key] (transposed A)
            _gemm_at(P, S, S * S, do, D, 64, dv, 3 * D, 64, S, 64, S, H)
            # dS = P * (dP - rowsum(dP*P)) / 8   (in place in dP)
            L.check(lib.mmvid_softmax_backward(_p(P), _p(dP), H * S, S, S, 0.125, _s()), "softmax_backward")
            # dQ = dS K            A = dS [S,S], B(k=key, n=d) = k[key, d]
            g(dP, S, 0, S * S, k, 1, 3 * D, 0, 64, dq, 3 * D, 0, 64, S, 64, S, 1, H)
            # dK = dS^T Q          A(m=key,k=q) = dS[q,key]
            _gemm_at(dP, S, S * S, q, 3 * D, 64, dk, 3 * D, 64, S, 64, S, H)
        return dqkv, None, None, None, None, None, None


def _gemm_at(A, lda, a_hstride, Bm, ldb, b_hstride, Cm, ldc, c_hstride, M, N, K, H):
    """C[h][m, n] = sum_k A[h][k, m] * B[h][k, n]  (A read transposed) via explicit per-head transposes."""
    # A[h] is [K, M] row-major with leading dim lda; transpose to [M, K] then use the regular batched GEMM
    lib = L.load()
    At = torch.empty(H, M, K, device=A.device, dtype=torch.float32)
    for h in range(H):
        src = A.view(-1)[h * a_hstride:h * a_hstride + K * lda].view(K, lda)[:, :M].contiguous()
        L.check(lib.mmvid_transpose2d(_p(src), _p(At[h]), K, M, _s()), "transpose2d")
    ops.gemm_batched_f32(At, K, 0, M * K, Bm, 1, ldb, 0, b_hstride, Cm, ldc, 0, c_hstride, M, N, K, 1, H)


class EmbedFn(torch.autograd.Function):
    """Fused embedding gather for ONE segment with autograd to (table, table2, pos)."""

    @staticmethod
    def forward(ctx, table, table2, pos, ids, B, S_total, seq_off, pad, out_buf):
        seg = dict(ids=ids, seq_off=seq_off, table=table.detach(), table2=None if table2 is None else table2.detach(),
                   pos=None if pos is None else pos.detach().contiguous(), pad=pad)
        ops.embed_gather(out_buf, [seg])
        ctx.seg = (ids, seq_off, pad, B, S_total)
        ctx.shapes = (table.shape, None if table2 is None else table2.shape, None if pos is None else pos.shape)
        ctx.mark_dirty(out_buf)
        return out_buf

    @staticmethod
    def backward(ctx, dx):
        lib = L.load()
        ids, seq_off, pad, B, S_total = ctx.seg
        ts, t2s, ps = ctx.shapes
        dx = dx.contiguous()
        D = dx.shape[-1]
        dt = torch.zeros(ts, device=dx.device, dtype=torch.float32)
        dt2 = torch.zeros(t2s, device=dx.device, dtype=torch.float32) if t2s is not None else None
        dp = torch.zeros(ps, device=dx.device, dtype=torch.float32) if ps is not None else None
        seg = L.EmbedSegment()
        seg.ids, seg.ids_bstride, seg.n, seg.seq_off = ids.data_ptr(), (ids.stride(0) if ids.shape[0] == B else 0), ids.shape[1], seq_off
        if pad is not None:
            seg.pad_value, seg.pad_base, seg.use_pad = pad[0], pad[1], 1
        seg.table_rows = ts[0]
        L.check(lib.mmvid_embed_backward(_p(dx), B, S_total, D, C.byref(seg), _p(dt), _p(dt2), _p(dp), _s()), "embed_backward")
        return dt, dt2, dp, None, None, None, None, None, None


def cross_entropy_selected(logits, target, sel):
    """mean over selected rows of CE(logits[row], target[row]) with gradient to logits (F.cross_entropy on
    logits[~mask1], dalle_bert.py:1040)."""
    return _CEFn.apply(logits, target, sel)


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, sel):
        lib = L.load()
        rows, n = logits.shape
        logits = logits.contiguous()
        dl = torch.empty_like(logits)
        loss_rows = torch.empty(rows, device=logits.device, dtype=torch.float32)
        sel_u8 = sel.to(torch.uint8).contiguous()
        L.check(lib.mmvid_cross_entropy(_p(logits), _p(target.contiguous()), _p(sel_u8), _p(dl), _p(loss_rows), rows, n, _s()),
                "cross_entropy")
        cnt = sel_u8.sum().clamp_min(1).float()
        ctx.save_for_backward(dl, cnt)
        return loss_rows.sum() / cnt

    @staticmethod
    def backward(ctx, g):
        dl, cnt = ctx.saved_tensors
        return dl * (g / cnt), None, None
