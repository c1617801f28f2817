"""Times one ART-V decode step (persistent cooperative kernel vs native per-layer launches) at fixed cache lengths."""
import argparse, sys, ctypes as C, torch
ap = argparse.ArgumentParser()
ap.add_argument("--impl", default="stream,fused,persistent,native")
ap.add_argument("--pos", default="400,1300,2300")
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--warm", type=int, default=3)
ARGS = ap.parse_args()
sys.path.insert(0, ".")
from mmvid_b200 import _lib as L, ops
lib = L.load()
B, D, H, NL, S_max = 4, 768, 12, 12, 2369
dev = "cuda"
g = torch.Generator().manual_seed(0)
def r(*s): return (torch.randn(*s, generator=g) * 0.02).to(dev)
layers = (L.DecodeLayer * NL)()
keep = []
for li in range(NL):
    t = dict(ln1_w=torch.ones(D, device=dev), ln1_b=torch.zeros(D, device=dev), in_w=r(3 * D, D), in_b=r(3 * D), out_w=r(D, D), out_b=r(D),
             ln2_w=torch.ones(D, device=dev), ln2_b=torch.zeros(D, device=dev), fc_w=r(4 * D, D), fc_b=r(4 * D), proj_w=r(D, 4 * D), proj_b=r(D),
             kcache=r(B, H, S_max, 64), vcache=r(B, H, S_max, 64))
    keep.append(t)
    for k, v in t.items(): setattr(layers[li], k, v.data_ptr())
ws = torch.zeros(int(lib.mmvid_artv_decode_workspace_floats(B, D, H)), device=dev)
# 16-bit copies for the streaming kernel (decode_stream.cu)
layers16 = (L.DecodeLayer16 * NL)()
keep16 = []
for li in range(NL):
    t = keep[li]
    t16 = {k: t[k].half() for k in ("in_w", "out_w", "fc_w", "proj_w", "kcache", "vcache")}
    keep16.append(t16)
    for k in ("ln1_w", "ln1_b", "in_b", "out_b", "ln2_w", "ln2_b", "fc_b", "proj_b"):
        setattr(layers16[li], k, t[k].data_ptr())
    for k, v in t16.items():
        setattr(layers16[li], k, v.data_ptr())
ws16 = torch.zeros(int(lib.mmvid_artv_decode_stream_workspace_floats(B, D, H)), device=dev)
head_w, head_b = r(1024, D), r(1024)
head_w16 = head_w.half()
lnw, lnb = torch.ones(D, device=dev), torch.zeros(D, device=dev)
logits = torch.empty(B, 1024, device=dev)
h0 = r(B, D) * 50
st = ops._stream()
for pos in [int(x) for x in ARGS.pos.split(",")]:
    for name in ARGS.impl.split(","):
        def call():
            h = h0.clone()
            if name == "stream":
                L.check(lib.mmvid_artv_decode_stream(layers16, NL, ops._ptr(h), ops._ptr(ws16), ops._ptr(lnw), ops._ptr(lnb),
                                                     ops._ptr(head_w16), ops._ptr(head_b), ops._ptr(logits), 1024, B, D, H, S_max,
                                                     pos, None, 1, st))
            elif name == "fused":
                L.check(lib.mmvid_artv_decode_fused(layers, NL, ops._ptr(h), ops._ptr(ws), ops._ptr(lnw), ops._ptr(lnb), ops._ptr(head_w),
                                                    ops._ptr(head_b), ops._ptr(logits), 1024, B, D, H, S_max, pos, st))
            elif name == "persistent":
                L.check(lib.mmvid_artv_decode_persistent(layers, NL, ops._ptr(h), ops._ptr(ws), ops._ptr(lnw), ops._ptr(lnb), ops._ptr(head_w),
                                                         ops._ptr(head_b), ops._ptr(logits), 1024, B, D, H, S_max, pos, st))
            else:
                L.check(lib.mmvid_artv_decode_step(layers, NL, ops._ptr(h), ops._ptr(ws), B, D, H, S_max, pos, st))
            return h
        for _ in range(ARGS.warm): call()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        import time
        e0.record()
        t0 = time.perf_counter()
        for _ in range(ARGS.reps): call()
        t_cpu = (time.perf_counter() - t0) / ARGS.reps * 1e6
        e1.record(); torch.cuda.synchronize()
        print(f"    (CPU enqueue {t_cpu:.0f} us per step)", end=" ")
        print(f"pos {pos} {name}: {e0.elapsed_time(e1) / ARGS.reps * 1000:.0f} us per step", flush=True)
    a = None
