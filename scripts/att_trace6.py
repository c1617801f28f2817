"""Extended timeline of CTA 0 of the persistent attention kernel (v6): MMA issue threads and softmax warps per global step."""
import os, sys
os.environ["MMVID_ATT_IMPL"] = "6"
os.environ["MMVID_ATT_TRACE_EXT"] = "1"
import torch
sys.path.insert(0, ".")
from mmvid_b200 import _lib as L, ops
from mmvid_b200._lib import MASK_PREV
lib = L.load()
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
B, S, H, D = 4, 2115, 12, 768
qkv = torch.randn(B * S, 3 * D, device="cuda")
odt = ops.act_dtype(prec)
for _ in range(2):
    ops.attention_tc(qkv, B, S, H, MASK_PREV, [65, 66], prec, out_dtype=odt)
buf = torch.zeros(4096, dtype=torch.int64, device="cuda")
L.check(lib.mmvid_debug_attention_trace(buf.data_ptr()))
ops.attention_tc(qkv, B, S, H, MASK_PREV, [65, 66], prec, out_dtype=odt)
torch.cuda.synchronize()
L.check(lib.mmvid_debug_attention_trace(None))
t = buf.cpu().tolist()
t0 = min(x for x in t if x > 0)
r = lambda i: (t[i] - t0) if t[i] else -1
print(f"{prec} v6 spin={os.environ.get('MMVID_ATT_SPIN', '0')} poly={os.environ.get('MMVID_ATT_POLY', '2')}")
print("step | MMA g: QK(t) issued, v_full passed, P(t) seen, PV(t) issued | softmax g warp0: loop top, S(t) seen, [s_free arrivals qd0..3], P(t) signalled")
for g in range(2):
    for st in list(range(0, 6)) + list(range(16, 22)) + list(range(30, 40)):
        mma = [r(1024 + g * 256 + st * 4 + k) for k in range(4)]
        sm = [r(2048 + g * 256 + st * 4 + k) for k in range(2)]
        sf = [r(3072 + (g * 4 + qd) * 64 + st) for qd in range(4)]
        sig = r(128 + g * 192 + st * 6 + 5) if st < 17 else -1
        print(f"{'AB'[g]} t={st:2d} | {mma} | {sm} {sf} sig {sig}")
