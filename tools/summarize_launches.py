#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time per kernel name over one training step.
usage: summarize_launches.py launches.csv [start_marker_kernel_substring]"""
import csv, sys, collections, re
path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "nsecond": 1, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1)
    rows.append((int(r["ID"]), r["Kernel Name"], ns))
# one step = from one gather_rows launch pair to the next adam_kernel
names = [n for _, n, _ in rows]
adv = [i for i, n in enumerate(names) if "adam_advance_kernel" in n]
adam = [i for i, n in enumerate(names) if "adam_kernel" in n]
if len(adv) >= 2:  # round 2: the device-side step counter's kernel opens every step
    lo, hi = adv[-2], adv[-1]
elif len(adam) >= 2:
    lo, hi = adam[0] + 1, adam[1] + 1
    # include row-adam launches that follow the dense adam
    while hi < len(names) and ("adam_rows" in names[hi] or "scatter_add" in names[hi]):
        hi += 1
    while lo < len(names) and ("adam_rows" in names[lo] or "scatter_add" in names[lo]):
        lo += 1
else:
    lo, hi = 0, len(rows)
step = rows[lo:hi]
tot = sum(ns for _, _, ns in step)
agg = collections.defaultdict(lambda: [0, 0.0])
def short(n):
    n = re.sub(r"\(.*", "", n)
    n = re.sub(r"savqa::\(anonymous namespace\)::", "", n)
    n = re.sub(r"void ", "", n)
    return n[:90]
for _, n, ns in step:
    a = agg[short(n)]
    a[0] += 1
    a[1] += ns
print(f"one step: launches {len(step)}  (IDs {step[0][0]}..{step[-1][0]}), serialized kernel time {tot/1e6:.3f} ms")
print(f"{'kernel':92s} {'n':>5s} {'ms':>9s} {'share':>7s} {'avg us':>9s}")
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:92s} {n:5d} {ns/1e6:9.3f} {100*ns/tot:6.1f}% {ns/n/1e3:9.2f}")
