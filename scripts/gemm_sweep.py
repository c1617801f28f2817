"""GEMM tuning sweep (run on the GPU box): times mmvid_linear at the benchmark shapes for forced tile widths /
rasterisations.  Each configuration runs in a fresh process because the overrides are read from the environment."""
import os, subprocess, sys, json
SHAPES = [(8460, 3072, 768, "c_fc"), (8460, 768, 3072, "c_proj"), (8460, 2304, 768, "qkv"), (8460, 768, 768, "out_proj"),
          (2115, 3072, 768, "c_fc_b1")]
CHILD = r'''
import sys, torch, json
sys.path.insert(0, ".")
from mmvid_b200 import ops
prec = sys.argv[1]
res = {}
flush = torch.empty(64*1024*1024, device="cuda")
for (M,N,K,name) in json.loads(sys.argv[2]):
    dt = torch.bfloat16 if prec == "bf16" else torch.float32
    a = torch.randn(M,K,device="cuda").to(dt); w = (torch.randn(N,K,device="cuda")/30).to(dt); b = torch.randn(N,device="cuda")
    out = torch.empty(M,N,device="cuda",dtype=dt)
    for _ in range(3): ops.linear(a,w,b,precision=prec,out=out)
    torch.cuda.synchronize(); tot=0
    for _ in range(10):
        flush.zero_(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear(a,w,b,precision=prec,out=out); e1.record(); torch.cuda.synchronize(); tot+=e0.elapsed_time(e1)
    # warm (no flush) back-to-back
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.linear(a,w,b,precision=prec,out=out)
    e1.record(); torch.cuda.synchronize(); warm=e0.elapsed_time(e1)/20
    res[name]=(round(tot/10*1000,1), round(2*M*N*K/(tot/10)/1e9,1), round(warm*1000,1), round(2*M*N*K/warm/1e9,1))
print(json.dumps(res))
'''
QUICK = "--quick" in sys.argv
if "--child" in sys.argv:   # one configuration in THIS process' environment: gemm_sweep.py --child <prec>
    prec = sys.argv[sys.argv.index("--child") + 1]
    r = subprocess.run([sys.executable, "-c", CHILD, prec, json.dumps(SHAPES)], capture_output=True, text=True, timeout=300)
    print(prec, r.stdout.strip() or r.stderr[-800:], flush=True)
    sys.exit(0)
if "--2cta" in sys.argv:
    for prec in ("tf32", "bf16"):
        for bn2 in (0, 128, 192, 256):
            env = dict(os.environ, MMVID_GEMM_2CTA=str(bn2))
            r = subprocess.run([sys.executable, "-c", CHILD, prec, json.dumps(SHAPES)], env=env, capture_output=True, text=True, timeout=300)
            print(prec, "2CTA BN", bn2, r.stdout.strip() or r.stderr[-800:], flush=True)
    sys.exit(0)
for prec in ("tf32", "bf16"):
    for bn in ((128, 256) if QUICK else (64, 128, 256)):
        for raster in ((0, 1) if QUICK else (0, 1, 2)):
            env = dict(os.environ, MMVID_GEMM_BN=str(bn), MMVID_GEMM_RASTER=str(raster))
            r = subprocess.run([sys.executable, "-c", CHILD, prec, json.dumps(SHAPES)], env=env, capture_output=True, text=True, timeout=300)
            print(prec, "BN", bn, "raster", raster, r.stdout.strip() or r.stderr[-500:], flush=True)
