"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and launch count per kernel family.

    python scripts/summarise_launches.py gpurun_out/launches.csv [> profiles/rN_launches.md]
"""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    v *= {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    rows.append((r["Kernel Name"], v))


def _clean(name):
    n = re.sub(r"^void ", "", name)
    return n.replace("<unnamed>::", "").replace("(anonymous namespace)::", "").replace("mmvid::", "")


def family(name):
    n = _clean(name)
    n = re.sub(r"<.*", "", n)
    return n.split("(")[0]


def variant(name):
    n = _clean(name).split("(")[0]
    m = re.search(r"<(.*)>", n)
    return m.group(1) if m else ""


tot = sum(v for _, v in rows)
fam = defaultdict(lambda: [0.0, 0])
var = defaultdict(lambda: [0.0, 0])
for n, v in rows:
    fam[family(n)][0] += v
    fam[family(n)][1] += 1
    var[(family(n), variant(n))][0] += v
    var[(family(n), variant(n))][1] += 1
print(f"{len(rows)} launches, {tot / 1000:.2f} ms of kernel time (cold-cache, serialised by ncu: compare SHARES)\n")
print("| kernel | launches | total ms | share | mean us |")
print("|---|---|---|---|---|")
for k, (t, c) in sorted(fam.items(), key=lambda kv: -kv[1][0]):
    print(f"| `{k}` | {c} | {t / 1000:.2f} | {100 * t / tot:.1f} % | {t / c:.1f} |")
print("\nTemplate instantiations with >= 1 % of the time:\n")
print("| kernel<...> | launches | total ms | share | mean us |")
print("|---|---|---|---|---|")
for (k, va), (t, c) in sorted(var.items(), key=lambda kv: -kv[1][0]):
    if t / tot >= 0.01:
        print(f"| `{k}<{va[:60]}>` | {c} | {t / 1000:.2f} | {100 * t / tot:.1f} % | {t / c:.1f} |")
