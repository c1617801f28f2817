"""Prints the pipeline timeline of CTA (0,0) of the flash-attention kernel at the benchmark shape (clock64 stamps)."""
import sys, ctypes as C, torch
sys.path.insert(0, ".")
from mmvid_b200 import _lib as L, ops
from mmvid_b200._lib import MASK_PREV
lib = L.load()
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32"
B, S, H, D = 4, 2115, 12, 768
qkv = torch.randn(B * S, 3 * D, device="cuda")
odt = torch.float32 if prec == "tf32" else torch.bfloat16
for _ in range(2):
    ops.attention_tc(qkv, B, S, H, MASK_PREV, [65, 66], prec, out_dtype=odt)
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
L.check(lib.mmvid_debug_attention_trace(buf.data_ptr()))
ops.attention_tc(qkv, B, S, H, MASK_PREV, [65, 66], prec, out_dtype=odt)
torch.cuda.synchronize()
L.check(lib.mmvid_debug_attention_trace(None))
t = buf.cpu().tolist()
n_kv = 17
t0 = min(x for x in t if x > 0)
print(f"{prec}: clock64 relative to first stamp; MMA warp: pA = p_ready_A seen, iA = PV_A(j)+QK_A(j+1) issued; softmax g: S ready, regs, max, exp, st landed, signalled")
for j in range(n_kv):
    m = [t[j * 4 + i] - t0 for i in range(4)]
    a = [t[128 + j * 6 + i] - t0 for i in range(6)]
    b = [t[128 + 192 + j * 6 + i] - t0 for i in range(6)]
    print(f"j={j:2d} MMA pA {m[0]:6d} iA {m[1]:6d} pB {m[2]:6d} iB {m[3]:6d} | A {a} | B {b}")
per = (t[(n_kv - 2) * 4] - t[2 * 4]) / (n_kv - 4)
print(f"steady-state period {per:.0f} clk per kv step")
