"""B200-native mirror of `mmvid_pytorch/transformers/clip_model.py:520-584` (OpenAICLIPTransformer) and of the
CLIP blocks it extracts (`ResidualAttentionBlock` :201-227, `Transformer` :230-247, `LayerNorm` :188, `QuickGELU` :196).

Same constructor arguments, same state-dict keys (`transformer.resblocks.N.{ln_1,ln_2,attn.in_proj_*,
attn.out_proj.*,mlp.c_fc.*,mlp.c_proj.*}`), same forward contract x[B,S,D] -> [B,S,D] in fp32.  torch.nn modules
are used purely as parameter containers; the compute is libmmvid_b200.so:

    per block:  LN -> QKV GEMM -> flash attention (analytic mask) -> out-proj GEMM (+residual)
                LN -> c_fc GEMM (+QuickGELU) -> c_proj GEMM (+residual)

precision: 'fp32' CUDA-core FFMA (bit-faithful parity path) | 'tf32' tcgen05 kind::tf32 (<=1e-3 parity mode)
           | 'bf16' tcgen05 kind::f16 (throughput mode; residual stream, LN, softmax and accumulators stay fp32)
"""
import os
from collections import OrderedDict

import torch
from torch import nn

from . import _lib, ops
from ._lib import ACT_QUICKGELU, BF16, FP32, H16, MASK_CAUSAL, MASK_NONE, MASK_PREV, TF32, PRECISIONS  # noqa: F401


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)  # parameter container only
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", nn.Identity()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)
        self.n_head = n_head


class Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.width, self.layers, self.heads = width, layers, heads
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads) for _ in range(layers)])


class _H16Cache:
    """bf16 / fp16 copies of fp32 parameters for the kind::f16 GEMMs, refreshed when a parameter is modified.
    (CLIP's Linear weights went through fp16 in the reference's own loader - clip_model.py:435-458 - so the fp16 copies
    of a CLIP-initialised transformer are exact.)"""

    def __init__(self):
        self._d = {}

    def get(self, p, dtype=torch.bfloat16):
        key = (id(p), dtype)
        tag = (p.data_ptr(), p._version, _lib.weights_epoch())
        hit = self._d.get(key)
        if hit is None or hit[0] != tag:
            w = p.detach()
            if dtype == torch.float16:
                w = w.clamp(-65504.0, 65504.0)
            hit = (tag, w.to(dtype).contiguous())
            self._d[key] = hit
        return hit[1]


class OpenAICLIPTransformer(nn.Module):
    def __init__(self, seq_len=0, which_model="openai_clip_text", model_path="ViT-B-32.pt", causal=True,
                 mask_type="causal", mask_kwargs={}, width=None, layers=None, precision="tf32"):
        super().__init__()
        self.context_length = seq_len
        self.causal = causal
        state = None
        if model_path is not None and os.path.exists(str(model_path)):
            # clip_model.py:535-543: the CLIP TorchScript archive; only one transformer stack is kept
            full = torch.jit.load(model_path, map_location="cpu").state_dict()
            prefix = "visual.transformer." if which_model == "openai_clip_visual" else "transformer."
            state = {k[len(prefix):]: v for k, v in full.items() if k.startswith(prefix)}
            width = state["resblocks.0.ln_1.weight"].shape[0]
            layers = len({k.split(".")[1] for k in state if k.startswith("resblocks.")})
        if which_model not in ("openai_clip_text", "openai_clip_visual"):
            raise NotImplementedError
        if width is None:
            width = 768 if which_model == "openai_clip_visual" else 512
        if layers is None:
            layers = 12
        heads = width // 64  # build_model: transformer_heads = width // 64 (clip_model.py:466,488)
        self.transformer = Transformer(width, layers, heads)
        if state is not None:
            # build_model casts Linear / MHA weights to fp16 (convert_weights :435-458) before .float() (:559)
            conv = {}
            for k, v in state.items():
                v = v.float()
                if ("in_proj" in k) or ("out_proj" in k) or ("c_fc" in k) or ("c_proj" in k):
                    v = v.half().float()
                conv[k] = v
            self.transformer.load_state_dict(conv)
        # mask (clip_model.py:545-578) kept analytically: no [S,S] tensor
        if causal:
            if mask_type == "causal":
                self.mask_kind, self.mask_rows = MASK_CAUSAL, ()
            elif mask_type == "mask_prev":
                self.mask_kind, self.mask_rows = MASK_PREV, tuple(int(i) for i in mask_kwargs["index"])
                if len(self.mask_rows) > 4:
                    raise NotImplementedError("at most 4 mask_prev rows")
            else:
                raise NotImplementedError
        else:
            self.mask_kind, self.mask_rows = MASK_NONE, ()
        self.precision = precision
        self._bf16 = _H16Cache()
        self._rows_dev = None

    # ------------------------------------------------------------------------------------------
    def _prev_rows_dev(self, device):
        if self._rows_dev is None or self._rows_dev.device != device:
            self._rows_dev = torch.tensor(list(self.mask_rows) or [0], dtype=torch.int32, device=device)
        return self._rows_dev

    def _qkv_buffers(self, B, S, H, prec, device):
        key = (B, S, H, prec, str(device))
        if getattr(self, "_qkv_key", None) != key:
            self._qkv_bufs = ops.alloc_qkv_buffers(B, H, S, prec, device)
            self._qkv_key = key
        return self._qkv_bufs

    def _w(self, p, prec):
        return self._bf16.get(p, ops.act_dtype(prec)) if prec in H16 else p.detach()

    @torch.no_grad()
    def forward(self, x, collect=None, kv_out=None, **kwargs):
        """x: float32 [B, S, D] (batch-first, like the reference wrapper :580-584) -> float32 [B, S, D]."""
        prec = PRECISIONS[self.precision] if isinstance(self.precision, str) else self.precision
        B, S, D = x.shape
        H = self.transformer.heads
        assert D == self.transformer.width
        x = x.contiguous().float().view(B * S, D)
        act_dt = ops.act_dtype(prec)
        first = True
        for li, blk in enumerate(self.transformer.resblocks):
            h = ops.layernorm(x, blk.ln_1.weight, blk.ln_1.bias, 1e-5, out_dtype=act_dt)
            if prec == FP32:
                qkv = ops.linear(h, self._w(blk.attn.in_proj_weight, prec), blk.attn.in_proj_bias, precision=prec)
                if kv_out is not None:  # prefill of the ART-V K/V cache: [B,H,S_max,64] per layer
                    kv = qkv.view(B, S, 3, H, 64)
                    kv_out[0][li][:, :, :S].copy_(kv[:, :, 1].permute(0, 2, 1, 3))
                    kv_out[1][li][:, :, :S].copy_(kv[:, :, 2].permute(0, 2, 1, 3))
                att = ops.attention_fp32(qkv, B, S, H, self.mask_kind, self._prev_rows_dev(x.device))
            else:
                # in-projection GEMM whose epilogue scatters Q, K and V^T straight into the attention layout
                bufs = self._qkv_buffers(B, S, H, prec, x.device)
                ops.linear_qkv(h, self._w(blk.attn.in_proj_weight, prec), blk.attn.in_proj_bias, bufs, B, S, H, prec)
                if kv_out is not None:
                    kv_out[0][li][:, :, :S].copy_(bufs[1][:, :, :S])
                    kv_out[1][li][:, :, :S].copy_(bufs[2][:, :, :, :S].transpose(2, 3))
                att = ops.attention_core(bufs, B, S, H, self.mask_kind, self.mask_rows, prec, out_dtype=act_dt)
            out = torch.empty_like(x) if first else x  # never write into the caller's tensor
            first = False
            x = ops.linear(att, self._w(blk.attn.out_proj.weight, prec), blk.attn.out_proj.bias, residual=x,
                           precision=prec, out=out)
            h = ops.layernorm(x, blk.ln_2.weight, blk.ln_2.bias, 1e-5, out_dtype=act_dt)
            h = ops.linear(h, self._w(blk.mlp.c_fc.weight, prec), blk.mlp.c_fc.bias, act=ACT_QUICKGELU, precision=prec,
                           out_dtype=act_dt)
            x = ops.linear(h, self._w(blk.mlp.c_proj.weight, prec), blk.mlp.c_proj.bias, residual=x, precision=prec,
                           out=x)
            if collect is not None:
                collect.append(x.view(B, S, D).clone())
        return x.view(B, S, D)
