"""Minimal reproduction harness for the CTA-pair fused QKV path: one linear_qkv call at shape A vs linear + split."""
import sys, torch
sys.path.insert(0, ".")
from mmvid_b200 import ops
B, S, H, D = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 2115, 12, 768
g = torch.Generator().manual_seed(0)
a = torch.randn(B * S, D, generator=g).cuda()
w = (torch.randn(3 * D, D, generator=g) / 28).cuda()
b = torch.randn(3 * D, generator=g).cuda()
bufs = ops.alloc_qkv_buffers(B, H, S, "tf32", a.device)
ops.linear_qkv(a, w, b, bufs, B, S, H, "tf32")
torch.cuda.synchronize()
qkv = ops.linear(a, w, b, precision="tf32").view(B, S, 3, H, 64)
q, k, vt = bufs
for name, got, ref in (("q", q[:, :, :S], qkv[:, :, 0].permute(0, 2, 1, 3)), ("k", k[:, :, :S], qkv[:, :, 1].permute(0, 2, 1, 3)),
                       ("vt", vt[:, :, :, :S], qkv[:, :, 2].permute(0, 2, 3, 1))):
    print(name, "max abs diff", float((got - ref).abs().max()))
