#!/usr/bin/env python
"""Per-source-line warp-stall samples of one ncu capture (page source): where a kernel's time goes.
usage: python scripts/ncu_lines.py gpurun_out/prof_X.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
r = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True)
cur = None; rows = []
hdr = None
for row in csv.reader(io.StringIO(r.stdout)):
    if len(row) == 2 and row[0] in ("File Path", "File Name"):
        cur = row[1].split("/")[-1]; continue
    if len(row) == 2:
        continue
    if row and row[0] == "Line No":
        hdr = row; continue
    if hdr is None or not row or not row[0].isdigit():
        continue
    si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
    try:
        rows.append((int(row[si] or 0), int(row[ii] or 0), cur, int(row[0]), row[1].strip()[:110]))
    except ValueError:
        pass
tot = sum(r[0] for r in rows) or 1
print(f"# {rep}: {tot} samples")
for s, i, f, ln, src in sorted(rows, reverse=True)[:top]:
    print(f"{100.0 * s / tot:5.1f}%  inst={i:8d}  {f}:{ln}  {src}")
