"""GPU parity of each C-ABI kernel against a plain PyTorch fp32 statement of the same op (run with -m gpu).

Tolerances: fp32 CUDA-core path ~1e-5 norm-relative (summation order only); TF32 tensor-core path 1e-3
(north_star's bar); FP16 path (kind::f16 with fp16 operands: the tf32 mantissa) 1e-3 as well; BF16 path 2e-2 (reported,
wide-range 16-bit mode).  Integer outputs must be bit-exact."""
import math

import pytest
import torch
import torch.nn.functional as F

from helpers import relerr

pytestmark = pytest.mark.gpu

TOL = {"fp32": 2e-5, "tf32": 1e-3, "bf16": 2e-2, "fp16": 1e-3}
H16 = {"bf16": torch.bfloat16, "fp16": torch.float16}


def _op(t, prec):
    """operand in the precision's own dtype (16-bit kinds take 16-bit operands)"""
    return t.to(H16[prec]) if prec in H16 else t


@pytest.fixture(scope="module", autouse=True)
def _fp32_reference_math():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def _ops():
    from mmvid_b200 import ops
    return ops


def test_library_loads_and_reports_version():
    from mmvid_b200 import _lib
    assert _lib.load().mmvid_version() >= 100


def test_embed_gather_segments():
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    B, S, D = 3, 11, 64
    table = torch.randn(40, D, generator=g).cuda()
    table2 = torch.randn(40, D, generator=g).cuda()
    pos = torch.randn(6, D, generator=g).cuda()
    ids = torch.randint(0, 30, (B, 6), generator=g).cuda()
    ids[:, 4:] = 0  # pads -> unique ids 34+i
    sp = torch.tensor([[1, 2]]).cuda()
    out = torch.zeros(B, S, D, device="cuda")
    ops.embed_gather(out, [dict(ids=sp, seq_off=0, table=table, table2=table2),
                           dict(ids=ids, seq_off=2, table=table, pos=pos, pad=(0, 34)),
                           dict(ids=ids[:, :3].contiguous(), seq_off=8, table=table2)])
    rid = torch.where(ids == 0, torch.arange(6, device="cuda") + 34, ids)
    ref = torch.cat([(table[sp] + table2[sp]).expand(B, 2, D), table[rid] + pos, table2[ids[:, :3]]], 1)
    assert torch.equal(out, ref)


def test_axial_table_matches_broadcast_sum():
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    D = 32
    w = [torch.randn(1, 3, 1, 1, D, generator=g).cuda(), torch.randn(1, 1, 4, 1, D, generator=g).cuda(),
         torch.randn(1, 1, 1, 5, D, generator=g).cuda()]
    t = ops.axial_table(w, (3, 4, 5))
    ref = ((w[0] + w[1]) + w[2]).reshape(60, D)
    assert torch.equal(t, ref)


@pytest.mark.parametrize("rows,D", [(5, 128), (1000, 768), (33, 1024)])
def test_layernorm(rows, D):
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    x = (torch.randn(rows, D, generator=g) * 3 + 1).cuda()
    w, b = torch.randn(D, generator=g).cuda(), torch.randn(D, generator=g).cuda()
    ref = F.layer_norm(x, (D,), w, b, 1e-5)
    assert relerr(ops.layernorm(x, w, b), ref) < 2e-6
    assert relerr(ops.layernorm(x, w, b, out_dtype=torch.bfloat16).float(), ref) < 5e-3


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16", "fp16"])
@pytest.mark.parametrize("M,N,K", [(1, 64, 64), (130, 192, 128), (2115, 2304, 768), (565, 768, 3072), (300, 1024, 768)])
def test_linear_bias_act_residual(prec, M, N, K):
    ops = _ops()
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    bias = torch.randn(N, generator=g).cuda()
    res = torch.randn(M, N, generator=g).cuda()
    ref0 = F.linear(a.double(), w.double(), bias.double())
    ref = (ref0 * torch.sigmoid(1.702 * ref0) + res.double()).float()
    out = ops.linear(_op(a, prec), _op(w, prec), bias, act=ops.ACT_QUICKGELU, residual=res, precision=prec)
    e = relerr(out, ref)
    assert e < TOL[prec], f"{prec} {M}x{N}x{K}: {e}"
    # plain (no epilogue) + in-place residual aliasing
    ref2 = (F.linear(a.double(), w.double()) + res.double()).float()
    buf = res.clone()
    ops.linear(_op(a, prec), _op(w, prec), None, residual=buf, precision=prec, out=buf)
    assert relerr(buf, ref2) < TOL[prec]


def test_linear_bf16_output_and_unaligned_n():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    a = torch.randn(200, 256, generator=g).cuda()
    w = (torch.randn(130, 256, generator=g) / 16).cuda()  # N not a multiple of the tile or of 4-aligned vectors? (130 % 4 = 2)
    ref = F.linear(a, w)
    out = ops.linear(a, w, precision="tf32")
    assert relerr(out, ref) < 1e-3
    o16 = ops.linear(a.bfloat16(), w[:128].bfloat16().contiguous(), precision="bf16", out_dtype=torch.bfloat16)
    assert o16.dtype == torch.bfloat16 and relerr(o16.float(), ref[:, :128]) < 2e-2
    h16 = ops.linear(a.half(), w[:128].half().contiguous(), precision="fp16", out_dtype=torch.float16)
    assert h16.dtype == torch.float16 and relerr(h16.float(), ref[:, :128]) < 1e-3


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("M,N,K,act", [(1400, 3072, 768, 1), (1300, 768, 768, 0), (1100, 1024, 768, 0), (1500, 2368, 768, 1),
                                       (300, 96, 256, 0), (700, 160, 128, 1)])
def test_16bit_results_through_tma_stores(prec, M, N, K, act):
    """16-bit outputs of every tile flavour (CTA-pair 256 / 192 / 128 wide, single-CTA 128 / 64 wide): 64-column groups
    leave as one {64 x 32} box, lone 32-column chunks (N = 2368 = 37 x 64, N = 96, N = 160) as {32 x 32} boxes; the M
    tail is clipped by the TMA unit.  fp16 stores saturate instead of producing inf."""
    ops = _ops()
    g = torch.Generator().manual_seed(M + N)
    a = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    bias = torch.randn(N, generator=g).cuda()
    ref = F.linear(a.double(), w.double(), bias.double())
    if act:
        ref = ref * torch.sigmoid(1.702 * ref)
    out = ops.linear(_op(a, prec), _op(w, prec), bias, act=act, precision=prec, out_dtype=H16[prec])
    assert out.dtype == H16[prec]
    # the 16-bit store rounds once more: half an ulp of an 11-bit (fp16) / 8-bit (bf16) significand
    assert relerr(out.float(), ref.float()) < (1.2e-3 if prec == "fp16" else 2e-2)
    if prec == "fp16":
        big = ops.linear(_op(a * 300, prec), _op(w * 300, prec), None, precision=prec, out_dtype=torch.float16)
        assert torch.isfinite(big.float()).all() and float(big.float().abs().max()) == 65504.0


def _attn_ref(qkv, B, S, H, mask):
    D = H * 64
    q, k, v = qkv.view(B, S, 3, H, 64).double().unbind(2)
    q, k, v = [t.transpose(1, 2) for t in (q, k, v)]
    att = q @ k.transpose(-1, -2) / 8.0
    if mask is not None:
        att = att + mask.double().to(att.device)
    o = torch.softmax(att, -1) @ v
    return o.transpose(1, 2).reshape(B * S, D).float()


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16", "fp16"])
@pytest.mark.parametrize("B,S,H,kind", [(2, 37, 2, "prev"), (1, 200, 3, "causal"), (1, 565, 12, "prev"), (2, 128, 2, "none"),
                                        (1, 300, 2, "causal")])
def test_attention_masks(prec, B, S, H, kind):
    ops = _ops()
    from oracle import mmvid_oracle as O
    g = torch.Generator().manual_seed(S + H)
    qkv = torch.randn(B * S, 3 * H * 64, generator=g).cuda()
    rows = (S // 3, S // 3 + 1)
    if kind == "prev":
        mask, mk, pr = O.build_attention_mask(S, "mask_prev", rows), ops.MASK_PREV, rows
    elif kind == "causal":
        mask, mk, pr = O.build_attention_mask(S, "causal"), ops.MASK_CAUSAL, ()
    else:
        mask, mk, pr = None, ops.MASK_NONE, ()
    ref = _attn_ref(qkv, B, S, H, mask)
    if prec == "fp32":
        out = ops.attention_fp32(qkv, B, S, H, mk, torch.tensor(list(pr) or [0], dtype=torch.int32, device="cuda"))
    else:
        out = ops.attention_tc(qkv, B, S, H, mk, pr, prec, out_dtype=H16.get(prec, torch.float32)).float()
    e = relerr(out, ref)
    assert e < TOL[prec] * (1.5 if prec == "fp16" else 1), f"{prec} {kind} S={S}: {e}"  # fp16: + the 16-bit output rounding


@pytest.mark.parametrize("prec", ["tf32", "bf16", "fp16"])
@pytest.mark.parametrize("kind", ["prev", "causal"])
def test_attention_large_score_range_exercises_lazy_rescale(prec, kind):
    """Scores with a wide dynamic range force the running-max reference to move many times (the O *= alpha path
    of the lazy-rescale schedule); growing key magnitude along the sequence makes later tiles dominate."""
    ops = _ops()
    from oracle import mmvid_oracle as O
    B, S, H = 1, 700, 2
    g = torch.Generator().manual_seed(99)
    qkv = torch.randn(B * S, 3 * H * 64, generator=g)
    ramp = torch.linspace(0.5, 4.0, S).repeat(B).unsqueeze(1)
    qkv[:, H * 64:2 * H * 64] *= ramp          # keys grow along the sequence -> max keeps increasing tile after tile
    qkv[:, :H * 64] *= 2.0
    qkv = qkv.cuda()
    if kind == "prev":
        mask, mk, pr = O.build_attention_mask(S, "mask_prev", (100, 101)), ops.MASK_PREV, (100, 101)
    else:
        mask, mk, pr = O.build_attention_mask(S, "causal"), ops.MASK_CAUSAL, ()
    ref = _attn_ref(qkv, B, S, H, mask)
    out = ops.attention_tc(qkv, B, S, H, mk, pr, prec, out_dtype=H16.get(prec, torch.float32)).float()
    e = relerr(out, ref)
    assert e < (3e-2 if prec == "bf16" else 3e-3), f"{prec} {kind}: {e}"


@pytest.mark.parametrize("prec", ["tf32", "bf16", "fp16"])
@pytest.mark.parametrize("B,S,H", [(2, 37, 2), (1, 565, 12), (3, 128, 4), (3, 515, 12), (2, 700, 12)])
def test_fused_qkv_projection_scatter_matches_split(prec, B, S, H):
    """mmvid_linear_qkv (GEMM epilogue writes Q,K,V^T in attention layout) vs linear + qkv_split."""
    ops = _ops()
    g = torch.Generator().manual_seed(B * 100 + S)
    D = H * 64
    x = torch.randn(B * S, D, generator=g).cuda()
    w = (torch.randn(3 * D, D, generator=g) / math.sqrt(D)).cuda()
    b = torch.randn(3 * D, generator=g).cuda()
    ref = F.linear(x.double(), w.double(), b.double()).float().view(B, S, 3, H, 64)
    bufs = ops.alloc_qkv_buffers(B, H, S, prec, "cuda")
    ops.linear_qkv(_op(x, prec), _op(w, prec), b, bufs, B, S, H, prec)
    q, k, vt = [t.float() for t in bufs]
    tol = TOL[prec] * (1.5 if prec == "fp16" else 1)
    assert relerr(q[:, :, :S], ref[:, :, 0].permute(0, 2, 1, 3)) < tol
    assert relerr(k[:, :, :S], ref[:, :, 1].permute(0, 2, 1, 3)) < tol
    assert relerr(vt[:, :, :, :S], ref[:, :, 2].permute(0, 2, 3, 1)) < tol
    assert float(q[:, :, S:].abs().max() if q.shape[2] > S else 0) == 0 and float(vt[:, :, :, S:].abs().max() if vt.shape[3] > S else 0) == 0


def test_vq_argmin_bit_exact_and_ties():
    ops = _ops()
    from oracle import mmvid_oracle as O
    g = torch.Generator().manual_seed(7)
    cb = (torch.randn(1024, 256, generator=g) * 0.3)
    z = torch.randn(777, 256, generator=g) * 0.2
    ref = torch.argmin(O.vq_distances(z, cb), dim=1)
    idx = ops.vq_argmin(z.cuda(), cb.cuda())
    assert idx.dtype == torch.int64 and torch.equal(idx.cpu(), ref)
    # exact ties: duplicated codewords -> lowest index must win (torch.argmin semantics)
    cb2 = cb.clone()
    cb2[900] = cb2[17]
    cb2[400] = cb2[17]
    z2 = cb2[[17, 400, 900, 5]] + 1e-4
    assert ops.vq_argmin(z2.cuda(), cb2.cuda()).cpu().tolist() == [17, 17, 17, 5]
    # empty input
    assert ops.vq_argmin(torch.empty(0, 256).cuda(), cb.cuda()).numel() == 0


def test_codebook_gather():
    ops = _ops()
    g = torch.Generator().manual_seed(8)
    cb = torch.randn(1024, 256, generator=g).cuda()
    ids = torch.randint(0, 1024, (3, 16), generator=g).cuda()
    assert torch.equal(ops.codebook_gather(ids, cb).view(3, 16, 256), cb[ids])


@pytest.mark.parametrize("C,HW", [(128, 64), (512, 16), (256, 1024)])
def test_groupnorm_swish(C, HW):
    ops = _ops()
    g = torch.Generator().manual_seed(C)
    N, H = 3, int(math.sqrt(HW))
    x = (torch.randn(N, C, H, H, generator=g) * 2 + 0.5).cuda()
    w, b = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
    ref = F.group_norm(x, 32, w, b, 1e-6)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    assert relerr(ops.groupnorm(x_nhwc, w, b).permute(0, 3, 1, 2), ref) < 5e-6
    assert relerr(ops.groupnorm(x_nhwc, w, b, swish=True).permute(0, 3, 1, 2), ref * torch.sigmoid(ref)) < 5e-6


def test_conv_out_fused_matches_groupnorm_swish_conv_clamp():
    """Decoder tail (model.py:578-581 + vae.py:55) in one kernel vs the torch composition."""
    ops = _ops()
    g = torch.Generator().manual_seed(21)
    for (N, C, H, W) in ((2, 128, 32, 32), (1, 128, 64, 64), (3, 128, 16, 48)):
        x = (torch.randn(N, C, H, W, generator=g) * 1.5 + 0.3).cuda()
        gw, gb = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
        w = (torch.randn(3, C, 3, 3, generator=g) / 10).cuda()
        b = torch.randn(3, generator=g).cuda()
        h = F.group_norm(x, 32, gw, gb, 1e-6)
        ref = (F.conv2d(h * torch.sigmoid(h), w, b, padding=1).clamp(-1, 1) + 1) * 0.5
        out = ops.conv_out_fused(x.permute(0, 2, 3, 1).contiguous(), gw, gb, w.permute(0, 2, 3, 1).contiguous(), b)
        assert out.shape == ref.shape and relerr(out, ref) < 1e-5


def _pack(w):
    return w.permute(0, 2, 3, 1).contiguous()


def test_conv_variants_match_torch():
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    N, Cin, Cout, H = 2, 128, 256, 16
    x = torch.randn(N, Cin, H, H, generator=g).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    xn = x.permute(0, 2, 3, 1).contiguous()
    res = torch.randn(N, H, H, Cout, generator=g).cuda()
    # 3x3 stride 1 pad 1 (+ residual)
    ref = F.conv2d(x, w, b, padding=1).permute(0, 2, 3, 1)
    assert relerr(ops.conv2d(xn, _pack(w), b), ref) < 1e-5
    assert relerr(ops.conv2d(xn, _pack(w), b, residual=res), ref + res) < 1e-5
    # Downsample: pad (0,1,0,1) then 3x3 stride 2 (model.py:77-81)
    ref = F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=2).permute(0, 2, 3, 1)
    assert relerr(ops.conv2d(xn, _pack(w), b, stride=2, pad=(0, 0), out_hw=(H // 2, H // 2)), ref) < 1e-5
    # Upsample: nearest x2 then 3x3 (model.py:56-62)
    ref = F.conv2d(F.interpolate(x, scale_factor=2.0, mode="nearest"), w, b, padding=1).permute(0, 2, 3, 1)
    assert relerr(ops.conv2d(xn, _pack(w), b, upsample=True), ref) < 1e-5
    # first layer: NCHW input in [0,1] with fused 2x-1; last layer: NCHW output with fused clamp/rescale
    img = torch.rand(N, 3, 32, 32, generator=g).cuda()
    w3 = (torch.randn(128, 3, 3, 3, generator=g) / 5).cuda()
    b3 = torch.randn(128, generator=g).cuda()
    ref = F.conv2d(2 * img - 1, w3, b3, padding=1).permute(0, 2, 3, 1)
    assert relerr(ops.conv2d(img, _pack(w3), b3, in_nchw=True, pre_affine=True), ref) < 1e-5
    wl = (torch.randn(3, Cin, 3, 3, generator=g) / 10).cuda()
    bl = torch.randn(3, generator=g).cuda()
    ref = (F.conv2d(x, wl, bl, padding=1).clamp(-1, 1) + 1) * 0.5
    assert relerr(ops.conv2d(xn, _pack(wl), bl, out_nchw=True, post_clamp=True), ref) < 1e-5


@pytest.mark.parametrize("N,Cin,Cout,H", [(2, 128, 128, 32), (3, 512, 512, 4), (1, 256, 128, 128), (5, 256, 512, 8),
                                          (2, 128, 256, 16)])
def test_conv3x3_tensor_core_tf32(N, Cin, Cout, H):
    """tcgen05 implicit-GEMM conv: 4-D TMA tiles with halo zero-fill vs F.conv2d (fp64 reference)."""
    ops = _ops()
    g = torch.Generator().manual_seed(N * 1000 + Cin + H)
    x = torch.randn(N, Cin, H, H, generator=g).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    res = torch.randn(N, H, H, Cout, generator=g).cuda()
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1).float()
    xn = x.permute(0, 2, 3, 1).contiguous()
    out = ops.conv2d(xn, _pack(w), b, precision="tf32")
    e = relerr(out, ref)
    assert e < 1e-3, e
    out = ops.conv2d(xn, _pack(w), b, residual=res, precision="tf32")
    assert relerr(out, ref + res) < 1e-3


@pytest.mark.parametrize("N,Cin,Cout,H", [(2, 128, 128, 32), (3, 512, 512, 4), (1, 256, 128, 128), (5, 256, 512, 8),
                                          (2, 512, 256, 16), (1, 128, 128, 256)])
def test_conv3x3_tensor_core_fp16_with_fp16_groupnorm_and_upsample_producers(N, Cin, Cout, H):
    """kind::f16 implicit-GEMM conv (fp16 NHWC activations through {64 ch, BW, BH, N} TMA boxes, fp16 packed weights, fp32
    accumulate / bias / residual / output) fed by the GroupNorm+swish and nearest-x2 kernels' fp16 outputs: each stage
    against fp64 at the tf32 tolerance (fp16 carries the same 10-bit mantissa)."""
    ops = _ops()
    g = torch.Generator().manual_seed(N * 1000 + Cin + H)
    x = torch.randn(N, Cin, H, H, generator=g).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    res = torch.randn(N, H, H, Cout, generator=g).cuda()
    gw, gb = torch.randn(Cin, generator=g).cuda(), torch.randn(Cin, generator=g).cuda()
    xn = x.permute(0, 2, 3, 1).contiguous()
    w16 = _pack(w).half()
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1).float()
    out = ops.conv2d(xn.half(), w16, b, precision="fp16")
    assert out.dtype == torch.float32 and relerr(out, ref) < 1e-3
    assert relerr(ops.conv2d(xn.half(), w16, b, residual=res, precision="fp16"), ref + res) < 1e-3
    # GroupNorm + swish -> fp16 -> conv
    gn = F.group_norm(x.double(), 32, gw.double(), gb.double(), 1e-6)
    gn = gn * torch.sigmoid(gn)
    t16 = ops.groupnorm(xn, gw, gb, swish=True, fast=True, out_dtype=torch.float16)
    assert t16.dtype == torch.float16 and relerr(t16.float(), gn.permute(0, 2, 3, 1).float()) < 6e-4
    ref2 = F.conv2d(gn, w.double(), b.double(), padding=1).permute(0, 2, 3, 1).float()
    assert relerr(ops.conv2d(t16, w16, b, precision="fp16"), ref2) < 1e-3
    if H <= 64:  # nearest x2 -> fp16 -> conv (Upsample, model.py:56-62)
        u16 = ops.upsample2x(xn, out_dtype=torch.float16)
        up = F.interpolate(x.double(), scale_factor=2.0, mode="nearest")
        assert u16.shape == (N, 2 * H, 2 * H, Cin) and relerr(u16.float(), up.permute(0, 2, 3, 1).float()) < 6e-4
        ref3 = F.conv2d(up, w.double(), b.double(), padding=1).permute(0, 2, 3, 1).float()
        assert relerr(ops.conv2d(u16, w16, b, precision="fp16"), ref3) < 1e-3


def test_upsample2x():
    ops = _ops()
    x = torch.randn(2, 5, 7, 8).cuda()
    ref = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(ops.upsample2x(x), ref)


def test_softmax_logits_and_noise():
    ops = _ops()
    g = torch.Generator().manual_seed(10)
    l = torch.randn(70, 1024, generator=g).cuda()
    nz = torch.randn(70, 1024, generator=g).cuda()
    assert relerr(ops.softmax_logits(l), torch.softmax(l, -1)) < 1e-6
    assert relerr(ops.softmax_logits(l, nz, 0.7), torch.softmax(l + 0.7 * nz, -1)) < 1e-6


def test_decode_kernels_match_full_attention():
    ops = _ops()
    g = torch.Generator().manual_seed(11)
    B, H, S_max, L = 3, 4, 50, 37
    D = H * 64
    qkv_all = torch.randn(B, L, 3 * D, generator=g).cuda()
    kc = torch.zeros(B, H, S_max, 64, device="cuda")
    vc = torch.zeros(B, H, S_max, 64, device="cuda")
    for t in range(L):
        ops.kv_append(qkv_all[:, t].contiguous(), kc, vc, t)
    out = ops.decode_attention(qkv_all[:, L - 1].contiguous(), kc, vc, L)
    ref = _attn_ref(qkv_all.reshape(B * L, 3 * D), B, L, H, torch.full((L, L), float("-inf")).triu_(1))
    assert relerr(out, ref.view(B, L, D)[:, -1]) < 1e-5
    a = torch.randn(B, 256, generator=g).cuda()
    w = torch.randn(100, 256, generator=g).cuda() / 16
    bias = torch.randn(100, generator=g).cuda()
    res = torch.randn(B, 100, generator=g).cuda()
    assert relerr(ops.linear_small_m(a, w, bias, residual=res), F.linear(a, w, bias) + res) < 1e-5


# ------------------------------------------------------------------------------------------------ round-1 additions
@pytest.mark.parametrize("bn2", ["128", "192", "256"])
@pytest.mark.parametrize("M,N,K,act,res", [(1400, 768, 3072, 0, True), (1300, 3072, 768, 1, False), (1100, 1024, 768, 0, False)])
def test_cta_pair_tiles_with_tma_store_epilogue_match_fp64(monkeypatch, bn2, M, N, K, act, res):
    """tc_gemm2.cu (cta_group::2 tiles 256 x 128 / 192 / 256, results through TMA bulk stores): bias, QuickGELU, residual and
    the clipped M / N tails against an fp64 statement; the same call on the single-CTA kernel must agree as well."""
    ops = _ops()
    g = torch.Generator().manual_seed(M + N + int(bn2))
    a = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    b = torch.randn(N, generator=g).cuda()
    r = torch.randn(M, N, generator=g).cuda() if res else None
    ref = F.linear(a.double(), w.double(), b.double())
    if act:
        ref = ref * torch.sigmoid(1.702 * ref)
    if res:
        ref = ref + r.double()
    monkeypatch.setenv("MMVID_GEMM_2CTA", bn2)
    out = ops.linear(a, w, b, act=act, residual=r, precision="tf32")
    monkeypatch.setenv("MMVID_GEMM_2CTA", "1")
    one = ops.linear(a, w, b, act=act, residual=r, precision="tf32")
    assert relerr(out, ref.float()) < TOL["tf32"]
    assert relerr(out, one) < 1e-5
    monkeypatch.setenv("MMVID_GEMM_TMA_STORE", "0")   # the transposing epilogue stays the fallback (bf16 outputs, odd N)
    monkeypatch.setenv("MMVID_GEMM_2CTA", bn2)
    old = ops.linear(a, w, b, act=act, residual=r, precision="tf32")
    assert relerr(old, out) < 1e-5


@pytest.mark.parametrize("B,S,H", [(2, 700, 12), (3, 515, 12), (1, 1200, 12)])
def test_fused_qkv_on_the_cta_pair_kernel_is_identical_to_the_single_cta_kernel(monkeypatch, B, S, H):
    """M >= 1024 takes the 256 x 256 pair tile (Q / K by 3-D TMA stores, V^T from registers, batch-straddling chunks row
    by row): same Q, K, V^T as the single-CTA scatter, zero padding untouched."""
    ops = _ops()
    g = torch.Generator().manual_seed(B * 1000 + S)
    D = H * 64
    x = torch.randn(B * S, D, generator=g).cuda()
    w = (torch.randn(3 * D, D, generator=g) / math.sqrt(D)).cuda()
    b = torch.randn(3 * D, generator=g).cuda()
    outs = []
    for pair in ("1", "0"):
        monkeypatch.setenv("MMVID_QKV_PAIR", pair)
        bufs = ops.alloc_qkv_buffers(B, H, S, "tf32", "cuda")
        ops.linear_qkv(x, w, b, bufs, B, S, H, "tf32")
        outs.append([t.clone() for t in bufs])
    ref = F.linear(x.double(), w.double(), b.double()).float().view(B, S, 3, H, 64)
    (q, k, vt), (q0, k0, vt0) = outs
    assert relerr(q[:, :, :S], ref[:, :, 0].permute(0, 2, 1, 3)) < TOL["tf32"]
    assert relerr(vt[:, :, :, :S], ref[:, :, 2].permute(0, 2, 3, 1)) < TOL["tf32"]
    for got, old in ((q, q0), (k, k0), (vt, vt0)):
        assert relerr(got, old) < 1e-6
    assert float(q[:, :, S:].abs().max()) == 0 and float(k[:, :, S:].abs().max()) == 0 and float(vt[:, :, :, S:].abs().max()) == 0


@pytest.mark.parametrize("prec", ["tf32", "bf16", "fp16"])
@pytest.mark.parametrize("poly", ["0", "2", "4"])
def test_attention_fma_pipe_exponentials_agree_with_fp64(monkeypatch, prec, poly):
    """The rotating-score-buffer kernel with 0 / 2 / 4 of every 8 exponentials on the FMA pipe (MMVID_ATT_POLY), both masks,
    causal items of different lengths, against the fp64 reference."""
    ops = _ops()
    from oracle import mmvid_oracle as O
    B, S, H = 6, 1100, 6
    g = torch.Generator().manual_seed(S)
    qkv = torch.randn(B * S, 3 * H * 64, generator=g).cuda()
    monkeypatch.setenv("MMVID_ATT_POLY", poly)
    odt = H16.get(prec, torch.float32)
    for kind, mk, rows in (("mask_prev", ops.MASK_PREV, (400, 401)), ("causal", ops.MASK_CAUSAL, ())):
        mask = O.build_attention_mask(S, kind, rows) if kind == "mask_prev" else O.build_attention_mask(S, "causal")
        ref = _attn_ref(qkv, B, S, H, mask)
        out = ops.attention_tc(qkv, B, S, H, mk, rows, prec, out_dtype=odt).float()
        assert relerr(out, ref) < TOL[prec] * (1.5 if prec == "fp16" else 1), f"{prec} poly {poly} {kind}"


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("spec", ["0", "1"])
def test_attention_speculative_exponentials_match_the_exact_max_path(monkeypatch, prec, spec):
    """MMVID_ATT_SPEC=1 (default in the 16-bit kinds): exponentials against the current reference without a row maximum,
    exact path only when a row sum exceeds 2^12.  Checked against fp64 on (a) ordinary scores, (b) scores whose maximum
    climbs tile after tile (the fallback fires on most steps), (c) rows that are fully masked in their first key tiles
    (mask_prev rows deep in the sequence: reference still -inf when the speculative pass runs), (d) a causal item."""
    ops = _ops()
    from oracle import mmvid_oracle as O
    monkeypatch.setenv("MMVID_ATT_SPEC", spec)
    odt = H16[prec]
    tol = TOL[prec] * (1.5 if prec == "fp16" else 1)
    B, S, H = 2, 900, 3
    g = torch.Generator().manual_seed(17)
    qkv = torch.randn(B * S, 3 * H * 64, generator=g)
    cases = [("plain", qkv.clone(), "mask_prev", (300, 301)), ("late_rows", qkv.clone(), "mask_prev", (700, 701)),
             ("causal", qkv.clone(), "causal", ())]
    ramp = qkv.clone()
    ramp[:, H * 64:2 * H * 64] *= torch.linspace(0.3, 5.0, S).repeat(B).unsqueeze(1)
    ramp[:, :H * 64] *= 2.0
    cases.append(("climbing_max", ramp, "mask_prev", (300, 301)))
    for name, x, kind, rows in cases:
        x = x.cuda()
        mask = O.build_attention_mask(S, kind, rows) if kind == "mask_prev" else O.build_attention_mask(S, "causal")
        ref = _attn_ref(x, B, S, H, mask)
        out = ops.attention_tc(x, B, S, H, ops.MASK_PREV if kind == "mask_prev" else ops.MASK_CAUSAL, rows, prec, out_dtype=odt).float()
        assert torch.isfinite(out).all(), name
        e = relerr(out, ref)
        assert e < tol * (3 if name == "climbing_max" else 1), f"{prec} spec={spec} {name}: {e}"


@pytest.mark.parametrize("prec", ["tf32", "fp16"])
@pytest.mark.parametrize("N,Cin,Cout,H", [(2, 128, 128, 32), (3, 256, 128, 16), (1, 128, 256, 64), (2, 512, 512, 16)])
def test_conv_epilogue_groupnorm_partials_match_a_statistics_pass(prec, N, Cin, Cout, H):
    """The tensor-core conv writes the GroupNorm partial statistics of its result (after bias and residual) from the
    epilogue; groupnorm(partial=) must give what the ordinary two-kernel statistics pass over the same result gives, and the
    conv result itself must not change."""
    ops = _ops()
    g = torch.Generator().manual_seed(N * 100 + Cout + H)
    x = (torch.randn(N, H, H, Cin, generator=g) * 1.5 + 0.3).cuda()
    w = (torch.randn(Cout, 3, 3, Cin, generator=g) / math.sqrt(Cin * 9)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    res = (torch.randn(N, H, H, Cout, generator=g) + 2.0).cuda()      # shifts the mean: sums are not centred in the epilogue
    gw, gb = torch.randn(Cout, generator=g).cuda(), torch.randn(Cout, generator=g).cuda()
    if prec == "fp16":
        x, w = x.half(), w.half()
    plain = ops.conv2d(x, w, b, residual=res, precision=prec)
    out, part = ops.conv2d(x, w, b, residual=res, precision=prec, gn_groups=32)
    assert part is not None, "these shapes are fusable"
    assert torch.equal(out, plain)
    for swish, dt in ((True, torch.float16), (False, torch.float32)):
        ref = ops.groupnorm(out, gw, gb, swish=swish, fast=True, out_dtype=dt).float()
        got = ops.groupnorm(out, gw, gb, swish=swish, fast=True, out_dtype=dt, partial=part).float()
        e = relerr(got, ref)
        assert e < (2e-3 if dt == torch.float16 else 2e-6), e   # fp16: one-ulp flips of the 16-bit rounding
    ref64 = F.group_norm(out.permute(0, 3, 1, 2).double(), 32, gw.double(), gb.double(), 1e-6).permute(0, 2, 3, 1).float()
    assert relerr(ops.groupnorm(out, gw, gb, partial=part), ref64) < 5e-6
    # not fusable: 8 x 8 images (two images per pixel tile) -> no partials, plain result
    xs = x[:, :8, :8].contiguous()
    o2, p2 = ops.conv2d(xs, w, b, precision=prec, gn_groups=32)
    assert p2 is None and torch.equal(o2, ops.conv2d(xs, w, b, precision=prec))


def test_groupnorm_streaming_kernel_exact_and_fast_swish():
    """groupnorm_apply2 (one channel quad per thread) for every VQGAN width, exact (expf / IEEE division) and MUFU swish."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    for C, HW in ((128, (24, 40)), (256, (16, 16)), (512, (8, 8))):
        x = (torch.randn(3, HW[0], HW[1], C, generator=g) * 2 + 0.5).cuda()
        w, b = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
        ref = F.group_norm(x.permute(0, 3, 1, 2).double(), 32, w.double(), b.double(), 1e-6)
        ref_s = (ref * torch.sigmoid(ref)).permute(0, 2, 3, 1).float()
        assert relerr(ops.groupnorm(x, w, b, swish=False), ref.permute(0, 2, 3, 1).float()) < 2e-5
        assert relerr(ops.groupnorm(x, w, b, swish=True), ref_s) < 2e-5
        assert relerr(ops.groupnorm(x, w, b, swish=True, fast=True), ref_s) < 2e-5


@pytest.mark.skipif(__import__("os").environ.get("MMVID_TEST_EXPERIMENTAL", "0") != "1",
                    reason="opt-in: code paths written after round 1's GPU budget ran out (MMVID_TEST_EXPERIMENTAL=1)")
@pytest.mark.parametrize("N,Cin,Cout,H", [(2, 128, 128, 32), (1, 256, 128, 128), (4, 128, 256, 16), (8, 512, 512, 8)])
def test_experimental_transposed_conv_tile_matches_fp64(monkeypatch, N, Cin, Cout, H):
    """MMVID_CONV_SWAP=1: 128 output channels x 256 pixels per tile (256-wide MMA), result chunks transposed before the
    bulk store.  Same reference and tolerance as the production conv test."""
    ops = _ops()
    g = torch.Generator().manual_seed(N * 1000 + Cin + H)
    x = torch.randn(N, Cin, H, H, generator=g).cuda()
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(Cin * 9)).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    res = torch.randn(N, H, H, Cout, generator=g).cuda()
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1).float()
    xn = x.permute(0, 2, 3, 1).contiguous()
    monkeypatch.setenv("MMVID_CONV_SWAP", "1")
    assert relerr(ops.conv2d(xn, _pack(w), b, precision="tf32"), ref) < 1e-3
    assert relerr(ops.conv2d(xn, _pack(w), b, residual=res, precision="tf32"), ref + res) < 1e-3


def test_mp_sample_draws_from_softmax_and_reports_the_drawn_probability():
    """mmvid_mp_sample (fused softmax + categorical draw + gather, dalle_bert.py:527-534): Y is exactly the softmax
    probability of the drawn token, the empirical distribution of 400 k draws of one row matches softmax (chi-square),
    rows flagged in `skip` are untouched, the draw is a function of (seed, offset) only, and Gumbel noise at temperature T
    changes the distribution the way the reference's formula says."""
    ops = _ops()
    g = torch.Generator().manual_seed(0)
    n = 1024
    base = (torch.randn(n, generator=g) * 2.0)
    base[5] = 6.0
    base[77] = -1e30  # zero-probability token
    R = 400_000
    logits = base.unsqueeze(0).repeat(R, 1).cuda().contiguous()
    Y = torch.zeros(R, device="cuda")
    tok = torch.full((R,), -1, dtype=torch.long, device="cuda")
    ops.mp_sample(logits, Y, tok, seed=1234, offset=0)
    p = torch.softmax(base.double(), 0)
    assert int(tok.min()) >= 0 and int(tok.max()) < n and not bool((tok == 77).any())
    assert torch.allclose(Y.cpu().double(), p[tok.cpu()], rtol=2e-5, atol=1e-9)
    counts = torch.bincount(tok.cpu(), minlength=n).double()
    big = p * R >= 20
    chi2 = float((((counts - p * R) ** 2) / (p * R))[big].sum())
    dof = int(big.sum()) - 1
    assert chi2 < dof + 6 * (2 * dof) ** 0.5, (chi2, dof)
    # determinism / seed sensitivity / skip
    Y2, tok2 = torch.zeros_like(Y), torch.zeros_like(tok)
    ops.mp_sample(logits, Y2, tok2, seed=1234, offset=0)
    assert torch.equal(tok, tok2) and torch.equal(Y, Y2)
    ops.mp_sample(logits, Y2, tok2, seed=1234, offset=1)
    assert not torch.equal(tok, tok2)
    skip = (torch.arange(R, device="cuda") % 3 == 0)
    Y3, tok3 = torch.full_like(Y, -7.0), torch.full_like(tok, -7)
    ops.mp_sample(logits, Y3, tok3, seed=99, offset=5, skip=skip)
    assert bool((tok3[skip] == -7).all()) and bool((Y3[skip] == -7.0).all()) and int(tok3[~skip].min()) >= 0
    # temperature: argmax of logits + T * gumbel follows softmax(logits / T); after the second softmax the draw is the
    # reference's two-stage procedure - compare against a torch restatement by Monte Carlo on a small alphabet
    small = torch.tensor([2.0, 1.0, 0.0, -1.0] + [-1e30] * 124).unsqueeze(0).repeat(200_000, 1).cuda().contiguous()
    Ys, ts = torch.zeros(200_000, device="cuda"), torch.zeros(200_000, dtype=torch.long, device="cuda")
    ops.mp_sample(small, Ys, ts, seed=7, offset=0, noise_scale=1.5)
    U = torch.rand(200_000, 4, generator=g).cuda()
    noisy = small[:, :4] + 1.5 * (-torch.log(-torch.log(U + 1e-20) + 1e-20))
    ref_tok = torch.multinomial(torch.softmax(noisy, 1), 1)[:, 0]
    f_ours = torch.bincount(ts.cpu(), minlength=4)[:4].double() / 200_000
    f_ref = torch.bincount(ref_tok.cpu(), minlength=4)[:4].double() / 200_000
    assert float((f_ours - f_ref).abs().max()) < 6e-3, (f_ours, f_ref)


def test_mp_keep_is_sampling_without_replacement_proportional_to_y():
    """mmvid_mp_keep (Gumbel-top-k): exactly k of the selectable tokens are kept (+ every preserved token), the next input
    ids are kept-token-or-[MASK], beams draw independently, and the inclusion frequencies match those of
    torch.multinomial(Y, k, replacement=False) (the call the reference makes, dalle_bert.py:651)."""
    ops = _ops()
    g = torch.Generator().manual_seed(1)
    Ttot, k, MASK = 96, 20, 1024
    y = torch.rand(Ttot, generator=g) ** 3 + 1e-4
    pm = torch.zeros(Ttot, dtype=torch.bool)
    pm[:8] = True
    trials = 20_000
    Y = y.unsqueeze(0).repeat(trials, 1).cuda().contiguous()
    I_tok = torch.randint(0, 1024, (trials, Ttot), generator=g).cuda()
    keep, ids_in = ops.mp_keep(Y, pm.cuda(), I_tok, k, MASK, seed=5, offset=0)
    assert keep.shape == (trials, Ttot) and keep.dtype == torch.bool
    assert bool(keep[:, :8].all()) and bool((keep[:, 8:].sum(1) == k).all())
    assert torch.equal(ids_in, torch.where(keep, I_tok, torch.full_like(I_tok, MASK)))
    ref = torch.multinomial(Y[:, 8:], k, replacement=False)
    ref_keep = torch.zeros(trials, Ttot - 8, dtype=torch.bool, device="cuda").scatter_(1, ref, True)
    f_ours, f_ref = keep[:, 8:].float().mean(0), ref_keep.float().mean(0)
    assert float((f_ours - f_ref).abs().max()) < 0.02, float((f_ours - f_ref).abs().max())
    # beams: independent draws from the same Y; fewer selectable tokens than k: all of them are kept
    keep2, _ = ops.mp_keep(Y[:4], pm.cuda(), I_tok[:4], k, MASK, seed=5, offset=1, beams=3)
    assert keep2.shape == (12, Ttot) and bool((keep2[:, 8:].sum(1) == k).all()) and not torch.equal(keep2[0], keep2[1])
    Yz = Y[:2].clone()
    Yz[:, 30:] = 0.0
    keep3, _ = ops.mp_keep(Yz, pm.cuda(), I_tok[:2], 40, MASK, seed=5, offset=2)
    assert bool(keep3[:, 8:30].all()) and not bool(keep3[:, 30:].any())
    # Shape-A sized rows (2048 tokens, k = 205 .. 1843)
    Yb = torch.rand(4, 2048, generator=g).cuda()
    for kk in (1, 205, 1843, 2048):
        kp, _ = ops.mp_keep(Yb, None, torch.zeros(4, 2048, dtype=torch.long, device="cuda"), kk, MASK, seed=1, offset=kk)
        assert bool((kp.sum(1) == kk).all())
