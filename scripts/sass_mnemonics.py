#!/usr/bin/env python
"""SASS mnemonic counts per kernel of libcps.so (cuobjdump -sass): the table under profiles/ that shows which kernels
use the TMA unit (UTMALDG / UTMASTG / UBLKCP), the FP64 tensor cores (DMMA), cp.async (LDGSTS), and that the LM kernels
issue DMUL / DADD and no DFMA outside the division / square-root expansions (the bit-exact contract).
Usage: python scripts/sass_mnemonics.py [libcps.so] > profiles/<tag>_sass_mnemonics.tsv"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                          "curvepointslam_b200", "libcps.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
cols = ["DMUL", "DADD", "DFMA", "DMMA", "LDGSTS", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF", "LDS", "STS", "BAR", "RED"]
counts, total, name = {}, collections.Counter(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and name:
        op = m.group(1)
        total[name] += 1
        for c in cols:
            if op == c or op.startswith(c + "."):
                counts[name][c] += 1
        if op in ("LDL", "STL") or op.startswith("LDL.") or op.startswith("STL."):
            counts[name]["SPILL"] += 1
print("# SASS mnemonic counts per kernel of libcps.so (cuobjdump -sass, sm_100a)")
print("# columns: kernel, instructions, " + ", ".join(cols) + ", LDL+STL (spills)")
for k in sorted(counts):
    print("\t".join([k, str(total[k])] + [str(counts[k][c]) for c in cols] + [str(counts[k]["SPILL"])]))
