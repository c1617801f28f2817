"""Timeline of CTA 0 of the single-CTA tcgen05 GEMM at the benchmark shapes (clock64 stamps, see gemm_stamp in tc_gemm.cu)."""
import os, sys
os.environ.setdefault("MMVID_GEMM_2CTA", "1")   # keep every shape on the single-CTA kernel (override: 128 | 192 | 256)
import torch
sys.path.insert(0, ".")
from mmvid_b200 import _lib as L, ops
lib = L.load()
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
dt = torch.float32 if prec == "tf32" else torch.bfloat16
for (M, N, K, name, act, res) in [(8460, 3072, 768, "c_fc", 1, False), (8460, 768, 3072, "c_proj", 0, True), (8460, 2304, 768, "qkv_plain", 0, False)]:
    a = torch.randn(M, K, device="cuda").to(dt); w = (torch.randn(N, K, device="cuda") / 30).to(dt); b = torch.randn(N, device="cuda")
    r = torch.randn(M, N, device="cuda") if res else None
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if res else dt)
    def run():
        ops.linear(a, w, b, act=act, residual=r, precision=prec, out=out)
    for _ in range(2): run()
    buf = torch.zeros(512, dtype=torch.int64, device="cuda")
    L.check(lib.mmvid_debug_gemm_trace(buf.data_ptr()))
    run(); torch.cuda.synchronize()
    L.check(lib.mmvid_debug_gemm_trace(None))
    t = buf.cpu().tolist()
    t0 = min(x for x in t if x > 0)
    nk = K // (32 if prec == "tf32" else 64)
    print(f"== {name} {prec} M={M} N={N} K={K} ({nk} k-blocks/tile; ideal MMA {nk * 256 if prec == 'tf32' else nk * 256} clk/tile)")
    for ti in range(8):
        g = lambda i: (t[ti * 64 + i] - t0) if t[ti * 64 + i] else -1
        kb = [t[ti * 64 + 24 + i] for i in range(min(nk, 32)) if t[ti * 64 + 24 + i]]
        d = [kb[i + 1] - kb[i] for i in range(len(kb) - 1)]
        print(f"tile {ti}: MMA acc_free {g(0)} kb0 {g(1)} kb_last {g(2)} committed {g(3)} | EPI full {g(8)} in_regs {g(9)} stored {g(10)} | "
              f"TMA first {g(16)} last {g(17)} | kb cadence min/med/max {min(d) if d else 0}/{sorted(d)[len(d)//2] if d else 0}/{max(d) if d else 0}")
