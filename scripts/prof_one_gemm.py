import sys, torch
sys.path.insert(0, ".")
from mmvid_b200 import ops
M, N, K = 8460, 3072, 768
a = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / 30; b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda")
for _ in range(3):
    ops.linear(a, w, b, precision="tf32", out=out)
torch.cuda.synchronize()
