#!/usr/bin/env python
"""Per-kernel SASS opcode census of lib/libsavqa_b200.so (cuobjdump -sass): the mnemonics that prove a Blackwell-native kernel
(B200_PROFILING.md) -- UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA, UTCBAR = tcgen05.commit -- and the
legacy tensor path (HMMA = mma.sync) that must NOT appear.  usage: python tools/sass_opcodes.py > profiles/r2_sass_opcodes.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "structured-alignment-vqa_b200", "lib", "libsavqa_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UBLKCP", "HMMA", "HGMMA", "SYNCS", "UCGABAR", "RED", "ATOM")
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        counts[cur]["_total"] += 1
        for w in WATCH:
            if op.startswith(w):
                counts[cur][w] += 1
demangled = subprocess.run(["cu++filt"] + list(counts), capture_output=True, text=True).stdout.splitlines() if counts else []
names = {k: (d if d else k) for k, d in zip(counts, demangled)}
print(f"# SASS opcode census of {os.path.relpath(lib, ROOT)} ({len(counts)} kernels); columns: " + " ".join(WATCH))
tot = collections.Counter()
for k, c in counts.items():
    n = re.sub(r"savqa::\(anonymous namespace\)::", "", names[k])
    n = n.replace("(int)", "").replace("(bool)", "").replace("void ", "").replace("savqa::<unnamed>::", "")
    n = re.sub(r"\(.*$", "", n)[:70]
    print(f"{n:70s} instr {c['_total']:6d}  " + "  ".join(f"{w}={c[w]}" for w in WATCH if c[w]))
    tot.update(c)
print("TOTAL " + "  ".join(f"{w}={tot[w]}" for w in WATCH))
assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor-core path found"
