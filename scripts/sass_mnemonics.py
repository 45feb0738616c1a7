#!/usr/bin/env python
"""Per-kernel counts of the Blackwell SASS mnemonics that prove which hardware paths a kernel uses (B200_PROFILING.md):
UTCHMMA / UTCQMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA load, .MULTICAST when multicast), UTMASTG / UTMAREDG (TMA
store / reduce), SYNCS (mbarrier), UCGABAR (cluster barrier).  Reads the in-tree libtq100.so with cuobjdump (no GPU needed).
    python scripts/sass_mnemonics.py > profiles/r02_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "snlp---tenary-post-train-quantization_b200", "libtq100.so")
PAT = ["UTCHMMA", "UTCQMMA", "LDTM", "UTMALDG", "MULTICAST", "UTMASTG", "UTMAREDG", "UBLKCP", "UBLKRED", "SYNCS", "UCGABAR", "LDGSTS",
       "REDUX", "MATCH"]
out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[kern] = collections.Counter()
        continue
    if kern:
        for p in PAT:
            if p in line:
                counts[kern][p] += 1
print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass; sm_100a)")
print(f"{'kernel':58s} " + " ".join(f"{p:>9s}" for p in PAT))
for k, c in counts.items():
    if sum(c.values()):
        print(f"{k[:58]:58s} " + " ".join(f"{c[p]:9d}" for p in PAT))
