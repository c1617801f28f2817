"""Timeline of CTA (0,0) of the attention kernel at the benchmark shape (the TRACE build: default POLY8, exact-max path), plus the
raw tcgen05.mma / MUFU / conversion / packed-FMA rates (mmvid_debug_mma_rate).  clock64 stamps relative to the first one."""
import os, sys, ctypes as C
os.environ.setdefault("MMVID_ATT_POLY", "2")
import torch
sys.path.insert(0, ".")
from mmvid_b200 import _lib as L, ops
from mmvid_b200._lib import MASK_PREV
lib = L.load()
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
out = torch.zeros(4, dtype=torch.int64, device="cuda")
names = ["SS tf32 128x128x8", "TS tf32 128x64x8", "SS f16 128x128x16", "TS f16 128x64x16", "SS tf32 128x256x8", "SS f16 128x256x16",
         "att pattern tf32 (16 TS + 8 SS)", "att pattern f16 (8 TS + 4 SS)"]
names += ["SS f16 128x64x16"]
for k, w in enumerate((1, 4, 8, 12)):
    L.check(lib.mmvid_debug_mma_rate(16 + k, 2000, out.data_ptr(), None))
    torch.cuda.synchronize()
    t = out.cpu().tolist()
    print(f"MUFU.EX2 {w:2d} warps: {t[0] / 16000:.2f} clk per warp instruction (warp 0), {t[1] / 16000:.2f} (last warp)")
for fl, nm in ((24, "tanh.approx"), (25, "rcp.approx")):
    L.check(lib.mmvid_debug_mma_rate(fl, 2000, out.data_ptr(), None))
    torch.cuda.synchronize()
    print(f"MUFU {nm}: {out.cpu().tolist()[0] / 16000:.2f} clk per warp instruction (one warp per sub-partition)")
for k, (nm, per) in enumerate((("8 x cvt.f16x2", 8), ("8 x ex2 + 4 x cvt.f16x2", 12), ("8 x fma.f32x2", 8), ("8 x ex2 + 8 x fma.f32x2", 16))):
    L.check(lib.mmvid_debug_mma_rate(20 + k, 2000, out.data_ptr(), None))
    torch.cuda.synchronize()
    t = out.cpu().tolist()
    print(f"PIPE {nm:26s}: {t[0] / 2000:.1f} clk per iteration = {t[0] / 2000 / per:.2f} clk per warp instruction (one warp per sub-partition)")
for fl in range(9):
    n = 960
    for _ in range(2):
        L.check(lib.mmvid_debug_mma_rate(fl, n, out.data_ptr(), None))
        torch.cuda.synchronize()
    t = out.cpu().tolist()
    print(f"MMA {names[fl]:34s}: issue {t[0] / n:6.1f} clk/mma, total {t[1] / n:6.1f} clk/mma")
B, S, H, D = 4, 2115, 12, 768
qkv = torch.randn(B * S, 3 * D, device="cuda")
odt = ops.act_dtype(prec)
for _ in range(2):
    ops.attention_tc(qkv, B, S, H, MASK_PREV, [65, 66], prec, out_dtype=odt)
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
L.check(lib.mmvid_debug_attention_trace(buf.data_ptr()))
ops.attention_tc(qkv, B, S, H, MASK_PREV, [65, 66], prec, out_dtype=odt)
torch.cuda.synchronize()
L.check(lib.mmvid_debug_attention_trace(None))
t = buf.cpu().tolist()
t0 = min(x for x in t if x > 0)
print(f"{prec} poly {os.environ.get('MMVID_ATT_POLY', '2')}: n = 2j+g; MMA: P(n) seen, PV(n)+QK(n+3) issued | softmax g(n) step j: S ready, regs, max, exps, st landed, signalled")
for n in range(0, 34):
    j, g = n >> 1, n & 1
    sm = [t[128 + g * 192 + j * 6 + i] - t0 for i in range(6)]
    if n >= 28 or n < 8 or 14 <= n < 20: print(f"n={n:2d} (j={j:2d} {'AB'[g]}) MMA seen {t[2 * n] - t0:6d} issued {t[2 * n + 1] - t0:6d} | softmax {sm}  busy {sm[5] - sm[0]}")

# arrival skew between the four warps of a softmax group (p_ready needs all four)
for g in range(2):
    for j in (6, 8, 10, 12):
        arr = [t[512 + (g * 4 + qd) * 32 + j] - t0 for qd in range(4)]
        print(f"tile {'AB'[g]} step {j}: p_ready arrivals of warps qd=0..3 {arr}  spread {max(arr) - min(arr)}; MMA saw it at {t[2 * (2 * j + g)] - t0}")
# steady-state period and phase durations over the middle key steps
for g in range(2):
    rows = [[t[128 + g * 192 + j * 6 + i] for i in range(6)] for j in range(4, 14)]
    per = [(rows[i + 1][0] - rows[i][0]) for i in range(len(rows) - 1)]
    ph = [[r[i + 1] - r[i] for i in range(5)] for r in rows]
    avg = [sum(p[i] for p in ph) / len(ph) for i in range(5)]
    print(f"tile {'AB'[g]}: period {sum(per) / len(per):.0f} clk; phases ld {avg[0]:.0f} max {avg[1]:.0f} exp+st issue {avg[2]:.0f} st wait {avg[3]:.0f} signal {avg[4]:.0f}"
          f" | wait for S {sum(rows[i + 1][0] - rows[i][5] for i in range(len(rows) - 1)) / (len(rows) - 1):.0f}")
