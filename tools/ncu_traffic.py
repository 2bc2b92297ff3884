#!/usr/bin/env python
"""Reads an `ncu --set full` report of tools/hbm_kernels.py and writes (i) a text summary of the DRAM / throughput counters per
kernel (profiles/r2_ncu_hbm_kernels.txt) and (ii) profiles/r2_ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per
launch, which bench.py reports as `roofline.traffic`.  usage: python tools/ncu_traffic.py gpurun_out/prof_hbm_r2.ncu-rep"""
import csv
import re
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size"]
ALGO = {"ln_fwd": 16384 * 512 * 18, "ln_bwd": 16384 * 512 * 14, "gather_rows": (1 << 18) * (2400 + 608 + 8), "colsum": 16384 * 2048 * 2,
        "gemm2": 16384 * 512 * 2 + 2048 * 512 * 2 + 16384 * 2048 * 2}


def num(r, k):
    try:
        v = float(r[col[k]].replace(",", ""))
    except Exception:
        return None
    u = units[col[k]].lower()
    if "byte" in u:
        v *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)
    if k == "gpu__time_duration.sum":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    return v


out_lines, traffic, seen = [], {}, {}
for r in data:
    name = r[col["Kernel Name"]]
    key = next((k for k in ALGO if k in name), None)
    if key is None:
        continue
    seen[key] = seen.get(key, 0) + 1
    if seen[key] != 2:  # second capture of each kernel (the first one also pays cold instruction caches)
        continue
    vals = {k: num(r, k) for k in want if k in col}
    rd, wr, us = vals.get("dram__bytes_read.sum"), vals.get("dram__bytes_write.sum"), vals.get("gpu__time_duration.sum")
    tot = (rd or 0) + (wr or 0)
    short = re.sub(r"^.*::", "", name.split("(")[0])
    out_lines.append(f"{short[:44]:44s} {us:8.1f} us  dram read {rd / 1e6:8.1f} MB  write {wr / 1e6:8.1f} MB  = {tot / us / 1e3:7.1f} GB/s under ncu  "
                     f"algorithmic {ALGO[key] / 1e6:7.1f} MB ({tot / ALGO[key]:.2f}x)  dram% {vals.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')}  "
                     f"L2% {vals.get('lts__throughput.avg.pct_of_peak_sustained_elapsed')}  sm% {vals.get('sm__throughput.avg.pct_of_peak_sustained_elapsed')}  "
                     f"tensor% {vals.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')}  regs {vals.get('launch__registers_per_thread')}")
    traffic[{"gemm2": "gemm2_conv1_M16384"}.get(key, key)] = {"dram_bytes": tot, "dram_read": rd, "dram_write": wr, "us_under_ncu": us,
                                                              "algorithmic_bytes": ALGO[key], "source": "profiles/r2_ncu_hbm_kernels.txt"}
hdr_txt = ("# round 2: ncu --set full --clock-control none of tools/hbm_kernels.py (L2 flushed before every launch; second capture of each kernel);\n"
           "# durations under ncu are serialised single launches; the CUDA-event figures are in the bench line's `hbm_kernels`\n")
open(os.path.join(ROOT, "profiles", "r2_ncu_hbm_kernels.txt"), "w").write(hdr_txt + "\n".join(out_lines) + "\n")
json.dump(traffic, open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json"), "w"), indent=1)
print(hdr_txt + "\n".join(out_lines))
