#!/usr/bin/env python
"""Key counters of an `ncu --set full` report (one line per metric, one column per captured launch), plus the stall
reasons summed over the source page.  usage: ncu_summary.py report.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.per_cycle_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_tensor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
for i, h in enumerate(hdr):
    if any(h == k or (k.endswith("cluster") and h.startswith(k)) or (k == "sm__inst_executed_pipe_tensor" and h.startswith(k) and h.endswith(".sum"))
           for k in keys):
        vals = [r[i] for r in data]
        if h == "Kernel Name":
            vals = [v[:90] for v in vals]
        print(f"{h} [{units[i]}]: {vals}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = None
agg = collections.Counter()
seen = set()
for r in rows:
    if "Address" in r and "Source" in r:
        hdr = r
        cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        ia = hdr.index("Address")
        continue
    if hdr is None or len(r) < len(hdr) or not r[ia].startswith("0x") or r[ia] in seen:
        continue
    seen.add(r[ia])
    for i in cols:
        if r[i] not in ("", "0"):
            agg[hdr[i][6:]] += int(r[i])
tot = sum(agg.values())
if tot:
    print("warp stall samples (first captured launch): " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in agg.most_common(9)))
