"""Library yardsticks at the benchmark shapes (NOT part of the product path): cuBLAS tf32 / bf16 GEMMs through torch.matmul
and torch SDPA (flash / cuDNN backends) at head dim 64, timed like bench.py (CUDA events, L2 flushed between launches).
Printed next to our own kernels' numbers in profiles/ to show how far each kernel is from what the vendor libraries reach
on the same silicon."""
import json
import torch
import torch.nn.functional as F

SHAPES = [(8460, 3072, 768, "c_fc"), (8460, 768, 3072, "c_proj"), (8460, 2304, 768, "qkv"), (8460, 768, 768, "out_proj")]
flush = torch.empty(64 * 1024 * 1024, device="cuda")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return sum(ts[: reps // 2]) / (reps // 2)


res = {}
for prec in ("tf32", "bf16"):
    torch.backends.cuda.matmul.allow_tf32 = True
    dt = torch.float32 if prec == "tf32" else torch.bfloat16
    for M, N, K, name in SHAPES:
        a = torch.randn(M, K, device="cuda").to(dt)
        w = (torch.randn(N, K, device="cuda") / 30).to(dt)
        out = torch.empty(M, N, device="cuda", dtype=dt)
        ms = timeit(lambda: torch.matmul(a, w.t(), out=out))
        res[f"cublas_{prec}_{name}"] = (round(ms * 1000, 1), round(2 * M * N * K / ms / 1e9, 1))
B, S, H = 4, 2115, 12
for dt, nm in ((torch.bfloat16, "bf16"),):
    q, k, v = [torch.randn(B, H, S, 64, device="cuda", dtype=dt) for _ in range(3)]
    for backend in ("flash", "cudnn"):
        try:
            from torch.nn.attention import sdpa_kernel, SDPBackend
            be = SDPBackend.FLASH_ATTENTION if backend == "flash" else SDPBackend.CUDNN_ATTENTION
            with sdpa_kernel(be):
                ms = timeit(lambda: F.scaled_dot_product_attention(q, k, v))
            res[f"sdpa_{backend}_{nm}"] = (round(ms * 1000, 1), round(4.0 * S * S * H * 64 * B / ms / 1e9, 1))
        except Exception as ex:  # backend not available for this build
            res[f"sdpa_{backend}_{nm}"] = f"unavailable: {str(ex)[:80]}"
print("YARD " + json.dumps(res))
