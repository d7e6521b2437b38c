#!/usr/bin/env python3
"""Stall samples / warp instructions of an .ncu-rep aggregated per function of a .cu file (by source line ranges).
usage: tools/ncu_funcs.py rep path/to/file.cu"""
import csv, io, subprocess, re, sys
rep, cu = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; hdr = None; per = {}
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and len(r) > 8 and r[2] == "-":
        d = dict(zip(hdr[4:], r[4:]))
        per[(cur, int(r[0]))] = (int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0))
src = open(cu).read().split("\n")
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r"^(?:template.*)?__(device|global)__.*?(\w+)\(", l)
    if m: funcs.append((i, m.group(2)))
funcs.append((len(src) + 1, "END"))
tot = sum(v[0] for v in per.values()); toti = sum(v[1] for v in per.values())
agg = {}
base = cu.split("/")[-1]
for (f, ln), (s, i) in per.items():
    if f == base:
        name = [n for (a, n), (b, _) in zip(funcs, funcs[1:]) if a <= ln < b]
        key = name[0] if name else "top"
    else:
        key = f
    a = agg.setdefault(key, [0, 0]); a[0] += s; a[1] += i
print(f"total samples {tot}, warp instructions {toti/1e6:.1f} M")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] * 1000 > tot or v[1] * 1000 > toti:
        print(f"{100*v[0]/tot:5.1f}% samples  {v[1]/1e6:9.1f} Minst  {k}")
