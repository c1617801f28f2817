"""Attention kernel variants at the benchmark shape: correctness vs an fp64 reference on the GPU + CUDA-event timing.

    python scripts/att_bench.py            # all variants, one subprocess each (a trapped kernel cannot poison the rest)
    python scripts/att_bench.py one <prec> <impl> <poly8> <spin> <dual>

Variants (argument order: prec impl poly spin dual spec; impl is kept for the log only):
MMVID_ATT_POLY = 0|2|4 (of every 8 exponentials on the FMA pipe), MMVID_ATT_SPIN = 0|1, MMVID_ATT_DUAL = 0|1.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def one(prec, impl, poly, spin, dual="1", spec="1"):
    os.environ["MMVID_ATT_POLY"], os.environ["MMVID_ATT_SPIN"] = poly, spin
    os.environ["MMVID_ATT_DUAL"] = dual
    os.environ["MMVID_ATT_SPEC"] = spec
    import torch
    from mmvid_b200 import ops
    from mmvid_b200._lib import MASK_PREV, MASK_CAUSAL
    odt = ops.act_dtype(prec)
    res = {"prec": prec, "impl": int(impl), "poly8": int(poly), "spin": int(spin), "dual": int(dual), "spec": int(spec)}

    def ref(qkv, B, S, H, kind, rows):
        q, k, v = qkv.view(B, S, 3, H, 64).double().unbind(2)
        q, k, v = [t.transpose(1, 2) for t in (q, k, v)]
        att = q @ k.transpose(-1, -2) / 8.0
        i = torch.arange(S, device=qkv.device)
        if kind == MASK_CAUSAL:
            att = att.masked_fill(i[None, :] > i[:, None], float("-inf"))
        else:
            for r in rows:
                att[:, :, r, :r] = float("-inf")
        return (torch.softmax(att, -1) @ v).transpose(1, 2).reshape(B * S, H * 64).float()

    errs = {}
    for name, (B, S, H, kind, rows) in {"prev_565": (1, 565, 12, MASK_PREV, (51, 52)), "causal_300": (1, 300, 2, MASK_CAUSAL, ()),
                                        "prev_2115": (1, 2115, 2, MASK_PREV, (65, 66)), "causal_2369": (1, 2369, 2, MASK_CAUSAL, ()),
                                        "prev_100": (2, 100, 2, MASK_PREV, (10, 11))}.items():
        g = torch.Generator().manual_seed(S)
        qkv = torch.randn(B * S, 3 * H * 64, generator=g).cuda()
        out = ops.attention_tc(qkv, B, S, H, kind, rows, prec, out_dtype=odt).float()
        r = ref(qkv, B, S, H, kind, rows)
        errs[name] = float((out - r).norm() / r.norm())
    # wide dynamic range: exercises the lazy rescale path
    B, S, H = 1, 700, 2
    g = torch.Generator().manual_seed(99)
    qkv = torch.randn(B * S, 3 * H * 64, generator=g)
    qkv[:, H * 64:2 * H * 64] *= torch.linspace(0.5, 4.0, S).unsqueeze(1)
    qkv[:, :H * 64] *= 2.0
    qkv = qkv.cuda()
    out = ops.attention_tc(qkv, B, S, H, MASK_CAUSAL, (), prec, out_dtype=odt).float()
    r = ref(qkv, B, S, H, MASK_CAUSAL, ())
    errs["rescale_700"] = float((out - r).norm() / r.norm())
    res["relerr"] = {k: round(v, 6) for k, v in errs.items()}

    # timing at shape A, batch 4 (kernel alone: Q/K/V^T prepared once), 256 MiB L2 flush between launches
    import ctypes as C
    from mmvid_b200 import _lib as L
    lib = L.load()
    B, S, H = 4, 2115, 12
    S_pad = (S + 127) // 128 * 128
    dt = ops.act_dtype(prec)
    qkv = torch.randn(B * S, 3 * H * 64, device="cuda")
    q = torch.zeros(B, H, S_pad, 64, device="cuda", dtype=dt)
    k = torch.zeros_like(q)
    vt = torch.zeros(B, H, 64, S_pad, device="cuda", dtype=dt)
    L.check(lib.mmvid_qkv_split(ops._ptr(qkv), ops._ptr(q), ops._ptr(k), ops._ptr(vt), ops._dt(q), B, H, S, S_pad, ops._stream()))
    o = torch.empty(B * S, H * 64, device="cuda", dtype=odt)
    pr = (C.c_int * 4)(65, 66, 0, 0)
    pid = L.PRECISIONS[prec]

    def att():
        L.check(lib.mmvid_attention(ops._ptr(q), ops._ptr(k), ops._ptr(vt), ops._ptr(o), ops._dt(o), o.stride(0), B, H, S, S_pad,
                                    MASK_PREV, pr, 2, pid, ops._stream()))
    flush = torch.empty(64 * 1024 * 1024, device="cuda", dtype=torch.float32)
    for _ in range(3):
        att()
    torch.cuda.synchronize()
    ts = []
    for _ in range(20):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        att()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = sum(ts[:10]) / 10
    res["us"] = round(ms * 1000, 1)
    res["tflops"] = round(4.0 * S * S * H * 64 * B / ms / 1e9, 1)
    if poly == "2" and (spec == "0" or prec == "tf32"):  # the timeline build: default poly8, exact-max path
        buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
        L.check(lib.mmvid_debug_attention_trace(buf.data_ptr()))
        att()
        torch.cuda.synchronize()
        L.check(lib.mmvid_debug_attention_trace(None))
        t = buf.cpu().tolist()
        res["period_clk"] = round((t[2 * 28] - t[2 * 8]) / 10.0)  # MMA rows: [2n] P(n) seen; period per key step = 2 tile-steps
        res["trace_softmaxA_j8"] = [t[128 + 8 * 6 + i] - t[128 + 8 * 6] for i in range(6)]
    print("ATT " + json.dumps(res), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        return one(*sys.argv[2:8])
    for prec in (sys.argv[1:] or ("tf32", "fp16", "bf16")):
        variants = [("5", "0", "0", "1", "0"), ("5", "2", "0", "1", "0")] if prec == "tf32" else \
                   [("5", "2", "0", "1", "0"), ("5", "2", "0", "1", "1"), ("5", "0", "0", "1", "0"), ("5", "4", "0", "1", "0")]
        for impl, poly, spin, dual, pp in variants:
            try:
                r = subprocess.run([sys.executable, __file__, "one", prec, impl, poly, spin, dual, pp], capture_output=True, text=True, timeout=100)
                lines = [l for l in r.stdout.splitlines() if l.startswith("ATT ")]
                print(lines[-1] if lines else f"ATT-FAIL {prec} impl={impl} poly={poly} spin={spin} dual={dual} rc={r.returncode}: {r.stderr[-600:]}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"ATT-TIMEOUT {prec} impl={impl} poly={poly} spin={spin}", flush=True)


if __name__ == "__main__":
    main()
