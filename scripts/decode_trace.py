"""Phase timeline of the streaming ART-V decode kernel (decode_stream.cu): %globaltimer stamps of CTA 0.

    python scripts/decode_trace.py [pos]
"""
import sys
import torch
sys.path.insert(0, ".")
from mmvid_b200 import _lib as L, ops
lib = L.load()
B, D, H, NL, S_max = 4, 768, 12, 12, 2369
pos = int(sys.argv[1]) if len(sys.argv) > 1 else 1300
dev = "cuda"
g = torch.Generator().manual_seed(0)
def r(*s): return (torch.randn(*s, generator=g) * 0.02).to(dev)
layers16 = (L.DecodeLayer16 * NL)()
keep = []
for li in range(NL):
    t = dict(ln1_w=torch.ones(D, device=dev), ln1_b=torch.zeros(D, device=dev), in_b=r(3 * D), out_b=r(D),
             ln2_w=torch.ones(D, device=dev), ln2_b=torch.zeros(D, device=dev), fc_b=r(4 * D), proj_b=r(D),
             in_w=r(3 * D, D).half(), out_w=r(D, D).half(), fc_w=r(4 * D, D).half(), proj_w=r(D, 4 * D).half(),
             kcache=r(B, H, S_max, 64).half(), vcache=r(B, H, S_max, 64).half())
    keep.append(t)
    for k, v in t.items():
        setattr(layers16[li], k, v.data_ptr())
ws = torch.zeros(int(lib.mmvid_artv_decode_stream_workspace_floats(B, D, H)), device=dev)
head_w, head_b = r(1024, D).half(), r(1024)
lnw, lnb = torch.ones(D, device=dev), torch.zeros(D, device=dev)
logits = torch.empty(B, 1024, device=dev)
h = r(B, D) * 50
st = ops._stream()
def step():
    L.check(lib.mmvid_artv_decode_stream(layers16, NL, ops._ptr(h), ops._ptr(ws), ops._ptr(lnw), ops._ptr(lnb), ops._ptr(head_w),
                                         ops._ptr(head_b), ops._ptr(logits), 1024, B, D, H, S_max, pos, None, 1, st))
for _ in range(3):
    step()
buf = torch.zeros(512, dtype=torch.int64, device=dev)
L.check(lib.mmvid_debug_decode_trace(buf.data_ptr()))
step()
torch.cuda.synchronize()
L.check(lib.mmvid_debug_decode_trace(None))
t = buf.cpu().view(64, 8)
t0 = int(t[0, 0])
names = ["QKV", "ATT", "OUT", "FC", "PROJ"]
print(f"pos {pos}: per phase (us): start | staged | LN | slab landed | items | stored | barrier passed   [duration]")
tot = {}
for ph in range(61):
    row = [int(x) for x in t[ph]]
    nm = names[ph % 5] if ph < 60 else "HEAD"
    rel = [(x - t0) / 1000.0 if x else float("nan") for x in row[:7]]
    end = row[6] if row[6] else row[5]
    dur = (end - row[0]) / 1000.0 if row[0] else float("nan")
    tot.setdefault(nm, []).append(dur)
    if ph < 10 or ph == 60:
        print(f"  {ph:2d} {nm:4s} " + " ".join(f"{x:8.2f}" for x in rel) + f"   [{dur:6.2f}]")
print("mean duration per phase kind (us):", {k: round(sum(v) / len(v), 2) for k, v in tot.items()})
print("token total (us):", (int(t[60, 5]) - t0) / 1000.0)
