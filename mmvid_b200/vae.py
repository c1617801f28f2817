"""B200-native mirror of `mmvid_pytorch/vae.py:15-71` (VQGanVAE1024) and of the taming VQGAN it wraps:
`taming/models/vqgan.py:16-75` (VQModel.encode/decode), `taming/modules/diffusionmodules/model.py`
(Encoder :363-466, Decoder :469-582, ResnetBlock :87-150, AttnBlock :153-205, Up/Downsample :45-84) and
`taming/modules/vqvae/quantize.py:230-341` (VectorQuantizer2).

Same public API (`get_codebook_indices`, `decode`, `decode_train`; attrs `num_layers=4`, `image_size`,
`num_tokens=1024`, `.model.quantize.embedding`), same state-dict keys (`model.encoder.down.L.block.B.conv1.weight`,
...).  torch.nn modules are parameter containers only; all compute is libmmvid_b200.so on NHWC fp32 activations:

  conv3x3 / strided / upsampled conv -> implicit GEMM (no im2col buffer); first conv reads NCHW and fuses `2x-1`
  (vae.py:41); last conv writes NCHW and fuses clamp(-1,1)*0.5+0.5 (vae.py:55); GroupNorm+swish fused;
  1x1 convs are GEMMs over pixel rows; AttnBlock = batched GEMM + row softmax; VQ = fused distance+argmin.
"""
import math
import os

import torch
from torch import nn

from . import _lib, ops
from ._lib import F16, FP32, MASK_NONE, PRECISIONS, TF32

# mmvid_pytorch/data/vqgan.1024.config.yml:5-21 (the only VQGAN configuration MMVID ships)
VQGAN_1024_CONFIG = dict(
    embed_dim=256, n_embed=1024,
    ddconfig=dict(double_z=False, z_channels=256, resolution=256, in_channels=3, out_ch=3, ch=128,
                  ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(16,), dropout=0.0))


def _norm(c):
    return nn.GroupNorm(num_groups=32, num_channels=c, eps=1e-6, affine=True)  # model.py:38-42


class ResnetBlock(nn.Module):
    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = _norm(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.norm2 = _norm(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, 1, 1, 0)


class AttnBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.in_channels = c
        self.norm = _norm(c)
        self.q = nn.Conv2d(c, c, 1)
        self.k = nn.Conv2d(c, c, 1)
        self.v = nn.Conv2d(c, c, 1)
        self.proj_out = nn.Conv2d(c, c, 1)


class _Resample(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 1, 1)  # Downsample uses stride 2 / pad 0 (model.py:71-75); shape is identical


class Encoder(nn.Module):
    def __init__(self, *, ch, ch_mult, num_res_blocks, attn_resolutions, in_channels, resolution, z_channels,
                 double_z=False, **ignore):
        super().__init__()
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.conv_in = nn.Conv2d(in_channels, ch, 3, 1, 1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = _Resample(block_in)
                curr_res //= 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, 3, 1, 1)


class Decoder(nn.Module):
    def __init__(self, *, ch, out_ch, ch_mult, num_res_blocks, attn_resolutions, in_channels, resolution, z_channels,
                 **ignore):
        super().__init__()
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, 3, 1, 1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = _Resample(block_in)
                curr_res *= 2
            self.up.insert(0, up)
        self.norm_out = _norm(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, 3, 1, 1)


class VectorQuantizer(nn.Module):
    def __init__(self, n_e, e_dim):
        super().__init__()
        self.n_e, self.e_dim = n_e, e_dim
        self.embedding = nn.Embedding(n_e, e_dim)
        self.embedding.weight.data.uniform_(-1.0 / n_e, 1.0 / n_e)  # quantize.py:254


class VQModel(nn.Module):
    def __init__(self, ddconfig, n_embed, embed_dim, **ignore):
        super().__init__()
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        self.quantize = VectorQuantizer(n_embed, embed_dim)
        self.quant_conv = nn.Conv2d(ddconfig["z_channels"], embed_dim, 1)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)


class _PackCache:
    """Conv weights repacked [Cout,Cin,KH,KW] -> [Cout,KH,KW,Cin] (K-contiguous for the implicit GEMM), cached."""

    def __init__(self):
        self._d = {}

    def conv(self, p, dtype=torch.float32):
        tag = (p.data_ptr(), p._version, _lib.weights_epoch())
        key = (id(p), dtype)
        hit = self._d.get(key)
        if hit is None or hit[0] != tag:
            w = p.detach().permute(0, 2, 3, 1)
            if dtype == torch.float16:
                w = w.clamp(-65504.0, 65504.0)
            hit = (tag, w.to(dtype).contiguous())
            self._d[key] = hit
        return hit[1]

    def cat(self, key, params):
        tag = tuple((p.data_ptr(), p._version) for p in params) + (_lib.weights_epoch(),)
        hit = self._d.get(key)
        if hit is None or hit[0] != tag:
            hit = (tag, torch.cat([p.detach().reshape(p.shape[0], -1) if p.dim() > 1 else p.detach() for p in params], 0)
                   .contiguous())
            self._d[key] = hit
        return hit[1]


class VQGanVAE1024(nn.Module):
    def __init__(self, vae_path=None, image_size=None, precision="fp32"):
        super().__init__()
        cfg = dict(VQGAN_1024_CONFIG)
        dd = dict(cfg["ddconfig"])
        if image_size:
            dd["resolution"] = image_size  # vae.py:24-25: decides which level carries AttnBlocks
        self.model = VQModel(dd, cfg["n_embed"], cfg["embed_dim"])
        if vae_path is not None:
            state = torch.load(vae_path, map_location="cpu")["state_dict"]
            self.model.load_state_dict(state, strict=False)
        self.num_layers = 4
        self.image_size = 256
        self.num_tokens = 1024
        self.precision = precision
        self._pack = _PackCache()

    # ------------------------------------------------------------------------------------------ building blocks
    def _prec(self):
        return PRECISIONS[self.precision] if isinstance(self.precision, str) else self.precision

    def _conv3(self, x, conv, **kw):
        """3x3 conv.  precision 'tf32' / 'fp16': stride-1 NHWC convs with Cin % 32 == 0 run on tcgen05 through 4-D TMA
        tiles (nearest-x2 upsampling is materialised first); the 3-channel input/output convs and the stride-2
        downsample stay on the fp32 implicit-GEMM path.  'fp16' (decoder): a float16 `x` (written by the GroupNorm /
        upsample kernel that precedes the conv) selects the kind::f16 implicit GEMM with fp16 packed weights."""
        # Decoder, tensor-core modes: every conv result is normalised next (Normalize, model.py:38-42), so the conv writes the
        # GroupNorm partial statistics of its result from the epilogue and the statistics pass over the activation (1 GB at
        # 256 px, batch 4 x 8 frames) never runs.  The pair (result, partials) waits in self._gn_slot for _gn().
        fuse = getattr(self, "_decoding", False) and os.environ.get("MMVID_GN_FUSE", "1") != "0"

        def run(x_, w_, prec):
            if not fuse or prec == FP32:
                return ops.conv2d(x_, w_, conv.bias, precision=prec, **kw)
            out, part = ops.conv2d(x_, w_, conv.bias, precision=prec, gn_groups=32, **kw)
            self._gn_slot = (out, part) if part is not None else None
            return out
        if x.dtype == torch.float16:
            w = self._pack.conv(conv.weight, torch.float16)
            assert kw.get("stride", 1) == 1 and not kw.get("upsample") and w.shape[3] % 64 == 0
            return run(x, w, F16)
        w = self._pack.conv(conv.weight)
        tc_ok = (self._prec() != FP32 and kw.get("stride", 1) == 1 and not kw.get("in_nchw") and not kw.get("out_nchw")
                 and w.shape[3] % 32 == 0 and w.shape[0] % 4 == 0)
        if tc_ok:
            if kw.pop("upsample", False):
                x = ops.upsample2x(x, out_dtype=self._act16())
                if x.dtype == torch.float16:
                    return run(x, self._pack.conv(conv.weight, torch.float16), F16)
            return run(x, w, TF32)
        return ops.conv2d(x, w, conv.bias, precision=FP32, **kw)

    def _gn(self, x, norm, **kw):
        """GroupNorm of x; uses the partial statistics its producing conv left in self._gn_slot (see _conv3)."""
        slot = getattr(self, "_gn_slot", None)
        self._gn_slot = None
        part = slot[1] if slot is not None and slot[0] is x else None
        return ops.groupnorm(x, norm.weight, norm.bias, partial=part, **kw)

    def _act16(self):
        """dtype the GroupNorm / upsample kernels hand to the 3x3 convs: float16 in the decoder of an 'fp16' VQGAN (same
        10-bit mantissa as the tf32 path at twice the MMA rate and half the bytes), float32 otherwise.  The encoder always
        stays on the fp32 / tf32 kernels: its VQ indices must not depend on the throughput mode."""
        return torch.float16 if (self._prec() == F16 and getattr(self, "_decoding", False)) else torch.float32

    def _conv1(self, x, conv, residual=None):
        N, H, W, C = x.shape
        w = conv.weight.detach().view(conv.weight.shape[0], C)
        out = ops.linear(x.view(-1, C), w, conv.bias, residual=None if residual is None else residual.view(-1, w.shape[0]),
                         precision=self._lin_prec())
        return out.view(N, H, W, w.shape[0])

    def _lin_prec(self):
        p = self._prec()
        return FP32 if p == FP32 else TF32  # 1x1 convs / spatial attention projections run TF32 at most

    def _resblock(self, x, blk):
        fast = self._prec() != FP32  # tensor-core modes: MUFU sigmoid; fp32 parity mode keeps expf + IEEE division
        a16 = self._act16()
        t = self._gn(x, blk.norm1, swish=True, fast=fast, out_dtype=a16)
        t = self._conv3(t, blk.conv1)
        t = self._gn(t, blk.norm2, swish=True, out=t, fast=fast, out_dtype=a16)
        sc = x if blk.in_channels == blk.out_channels else self._conv1(x, blk.nin_shortcut)
        return self._conv3(t, blk.conv2, residual=sc)

    def _attnblock(self, x, blk):
        N, H, W, C = x.shape
        HW = H * W
        t = self._gn(x, blk.norm, swish=False)
        wqkv = self._pack.cat(("qkv", id(blk)), [blk.q.weight, blk.k.weight, blk.v.weight])
        bqkv = self._pack.cat(("bqkv", id(blk)), [blk.q.bias, blk.k.bias, blk.v.bias])
        qkv = ops.linear(t.view(-1, C), wqkv, bqkv, precision=self._lin_prec())  # [N*HW, 3C]
        scores = torch.empty(N, HW, HW, device=x.device, dtype=torch.float32)
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
        # w_[b,i,j] = sum_c q[b,i,c] k[b,j,c] * C^-0.5 ; softmax over j ; h[b,i,c] = sum_j w_[b,i,j] v[b,j,c]
        ops.gemm_batched_f32(q, 3 * C, HW * 3 * C, 0, k, 3 * C, 1, HW * 3 * C, 0, scores, HW, HW * HW, 0, HW, HW, C, N, 1,
                             alpha=float(int(C) ** (-0.5)))
        ops.softmax_rows(scores, HW, HW, HW, N, MASK_NONE)
        att = torch.empty(N * HW, C, device=x.device, dtype=torch.float32)
        ops.gemm_batched_f32(scores, HW, HW * HW, 0, v, 1, 3 * C, HW * 3 * C, 0, att, C, HW * C, 0, HW, C, HW, N, 1)
        wp = blk.proj_out.weight.detach().view(C, C)
        out = ops.linear(att, wp, blk.proj_out.bias, residual=x.view(-1, C), precision=self._lin_prec())
        return out.view(N, H, W, C)

    # ------------------------------------------------------------------------------------------ encoder / decoder
    @torch.no_grad()
    def _encode_prequant(self, img):
        """img float32 [N,3,H,W] in [0,1] -> pre-quantisation latent rows [N, h, w, 256] (NHWC)."""
        enc = self.model.encoder
        h = self._conv3(img.contiguous().float(), enc.conv_in, in_nchw=True, pre_affine=True)
        for lvl in range(enc.num_resolutions):
            down = enc.down[lvl]
            for b in range(enc.num_res_blocks):
                h = self._resblock(h, down.block[b])
                if len(down.attn) > 0:
                    h = self._attnblock(h, down.attn[b])
            if lvl != enc.num_resolutions - 1:
                N, H, W, C = h.shape
                h = self._conv3(h, down.downsample.conv, stride=2, pad=(0, 0), out_hw=(H // 2, W // 2))  # model.py:77-81
        h = self._resblock(h, enc.mid.block_1)
        h = self._attnblock(h, enc.mid.attn_1)
        h = self._resblock(h, enc.mid.block_2)
        h = ops.groupnorm(h, enc.norm_out.weight, enc.norm_out.bias, swish=True, out=h)
        h = self._conv3(h, enc.conv_out)
        return self._conv1(h, self.model.quant_conv)

    @torch.no_grad()
    def get_codebook_indices(self, img):
        """vae.py:38-43: float [N,3,H,W] in [0,1] -> long [N, (H/16)^2]."""
        b = img.shape[0]
        z = self._encode_prequant(img)
        idx = ops.vq_argmin(z.view(-1, z.shape[-1]), self.model.quantize.embedding.weight.detach())
        return idx.view(b, -1)

    @torch.no_grad()
    def _decode_latent(self, z):
        """z: float32 NHWC [N, h, w, 256] codebook vectors -> float [N,3,H,W] in [0,1]."""
        self._decoding = True
        self._gn_slot = None
        try:
            return self._decode_latent_impl(z)
        finally:
            self._decoding = False
            self._gn_slot = None

    def _decode_latent_impl(self, z):
        dec = self.model.decoder
        h = self._conv1(z, self.model.post_quant_conv)
        h = self._conv3(h, dec.conv_in)
        h = self._resblock(h, dec.mid.block_1)
        h = self._attnblock(h, dec.mid.attn_1)
        h = self._resblock(h, dec.mid.block_2)
        for lvl in reversed(range(dec.num_resolutions)):
            up = dec.up[lvl]
            for bi in range(dec.num_res_blocks + 1):
                h = self._resblock(h, up.block[bi])
                if len(up.attn) > 0:
                    h = self._attnblock(h, up.attn[bi])
            if lvl != 0:
                h = self._conv3(h, up.upsample.conv, upsample=True)  # nearest x2 folded into the gather (model.py:56-62)
        # norm_out -> swish -> conv_out -> clamp/rescale in one pass over the 128-channel activation
        slot, self._gn_slot = getattr(self, "_gn_slot", None), None
        return ops.conv_out_fused(h, dec.norm_out.weight, dec.norm_out.bias, self._pack.conv(dec.conv_out.weight),
                                  dec.conv_out.bias, post_clamp=True, fast=self._prec() != FP32,
                                  partial=slot[1] if slot is not None and slot[0] is h else None)

    @torch.no_grad()
    def decode(self, img_seq):
        """vae.py:45-56: long [N, n] -> float [N,3,H,W] in [0,1]."""
        b, n = img_seq.shape
        hw = int(math.sqrt(n))
        z = ops.codebook_gather(img_seq, self.model.quantize.embedding.weight.detach()).view(b, hw, hw, -1)
        return self._decode_latent(z)

    @torch.no_grad()
    def decode_train(self, probs):
        """vae.py:58-68 (soft one-hot decode): probs [B, n, n_codes] -> images (forward only)."""
        b, n, d = probs.shape
        hw = int(math.sqrt(n))
        cb_t = self._pack.cat(("cb_t", id(self)), [self.model.quantize.embedding.weight]).t().contiguous()
        z = ops.linear(probs.reshape(-1, d).float().contiguous(), cb_t, precision=FP32)
        return self._decode_latent(z.view(b, hw, hw, -1))

    def forward(self, img):
        raise NotImplementedError
