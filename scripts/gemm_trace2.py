"""Timeline of cluster 0's leader CTA of the CTA-pair tcgen05 GEMM (tc_gemm2.cu, g2_stamp) at the benchmark shapes.

    python scripts/gemm_trace2.py [fp16|bf16|tf32]
"""
import sys
import torch
sys.path.insert(0, ".")
from mmvid_b200 import _lib as L, ops
lib = L.load()
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
dt = ops.act_dtype(prec)
M = 8460
for (N, K, name, act, res, odt) in [(3072, 768, "c_fc", 1, False, dt), (3072, 768, "c_fc_noact", 0, False, dt), (768, 3072, "c_proj", 0, True, torch.float32),
                                    (768, 768, "out_proj", 0, True, torch.float32), (1024, 768, "head", 0, False, torch.float32)]:
    a = torch.randn(M, K, device="cuda").to(dt)
    w = (torch.randn(N, K, device="cuda") / 30).to(dt)
    b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda") if res else None
    out = torch.empty(M, N, device="cuda", dtype=odt)
    tile = lib.mmvid_debug_pick_tile(M, N, K, L.PRECISIONS[prec], ops._dt(out))

    def run():
        ops.linear(a, w, b, act=act, residual=r, precision=prec, out=out)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    buf = torch.zeros(512, dtype=torch.int64, device="cuda")
    L.check(lib.mmvid_debug_gemm_trace(buf.data_ptr()))
    run()
    torch.cuda.synchronize()
    L.check(lib.mmvid_debug_gemm_trace(None))
    t = buf.cpu().tolist()
    t0 = min(x for x in t if x > 0)
    nk = K // (32 if prec == "tf32" else 64)
    print(f"== {name} {prec} M={M} N={N} K={K}: tile code {tile}, {nk} k-blocks/tile, {us:.1f} us warm (L2-hot), "
          f"{2.0 * M * N * K / us / 1e6:.0f} TFLOP/s")
    for ti in range(7):
        g = lambda i: (t[ti * 64 + i] - t0) if t[ti * 64 + i] else -1
        kb = [t[ti * 64 + 24 + i] for i in range(min(nk, 32)) if t[ti * 64 + 24 + i]]
        d = [kb[i + 1] - kb[i] for i in range(len(kb) - 1)]
        print(f"  tile {ti}: MMA acc_free {g(0)} kb0 {g(1)} kb_last {g(2)} committed {g(3)} | EPI full {g(8)} in_regs {g(9)} stored {g(10)}"
              f" | k-block cadence min/med/max {min(d) if d else 0}/{sorted(d)[len(d) // 2] if d else 0}/{max(d) if d else 0}")
