#!/usr/bin/env python3
"""Per-CUDA-source-line stall samples / instructions from an .ncu-rep (needs -lineinfo and --import-source on).
usage: tools/ncu_lines.py rep [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and r[0].isdigit() and len(r) > 8 and r[2] == "-":
        d = dict(zip(hdr[4:], r[4:]))
        out.append((int(d.get("# Samples", 0) or 0), int(d.get("Instructions Executed", 0) or 0), cur, int(r[0]), r[1].strip()[:110]))
tot = sum(o[0] for o in out) or 1; toti = sum(o[1] for o in out) or 1
print(f"total samples {tot}, total warp instructions {toti}")
for s, i, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{100*s/tot:5.1f}% smp {100*i/toti:5.1f}% inst  {f}:{ln}: {src}")
