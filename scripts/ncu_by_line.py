#!/usr/bin/env python
"""Aggregates an ncu report's warp-sampling data per CUDA source line / line range.
usage: ncu_by_line.py report.ncu-rep [file:lo-hi=name ...]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    loc, name = a.split("=")
    f, rng = loc.split(":")
    lo, hi = map(int, rng.split("-"))
    regions.append((f, lo, hi, name))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
fname, hdr = None, None
per_line = collections.Counter()
per_line_inst = collections.Counter()
src_text = {}
for r in csv.reader(txt.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Function Name"):
        try:
            line = int(r[0])
            samples = int(r[hdr.index("# Samples")])
            inst = int(r[hdr.index("Instructions Executed")])
        except ValueError:
            continue
        per_line[(fname, line)] += samples
        per_line_inst[(fname, line)] += inst
        src_text[(fname, line)] = r[1][:100]
tot = sum(per_line.values()) or 1
toti = sum(per_line_inst.values()) or 1
print("total samples", tot, "warp instructions", toti)
agg = collections.Counter()
aggi = collections.Counter()
for (f, l), s in per_line.items():
    name = f
    for rf, lo, hi, nm in regions:
        if rf == f and lo <= l <= hi:
            name = nm
            break
    agg[name] += s
    aggi[name] += per_line_inst[(f, l)]
for k, v in agg.most_common():
    print("%-28s samples %5.1f%%   instr %5.1f%%" % (k, 100 * v / tot, 100 * aggi[k] / toti))
print("--- top lines")
for (f, l), s in per_line.most_common(16):
    print("%5.1f%% inst %5.1f%%  %s:%d  %s" % (100 * s / tot, 100 * per_line_inst[(f, l)] / toti, f, l, src_text[(f, l)]))
