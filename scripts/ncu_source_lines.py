#!/usr/bin/env python
"""Per-source-line totals of one ncu capture (needs -lineinfo and --import-source on):
usage: python scripts/ncu_source_lines.py gpurun_out/prof_X.ncu-rep [top_n] [kernel regex (multi-kernel captures)]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
kern = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + kern, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur, hdr, lines = None, None, []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        d = dict(zip(hdr, r)); d["Source"] = r[1]
        lines.append((cur, int(r[0]), d))
tot = sum(int(d["# Samples"]) for _, _, d in lines) or 1
toti = sum(int(d["Instructions Executed"]) for _, _, d in lines) or 1
print(f"# {rep}: {tot} stall samples, {toti} warp instructions; top lines by samples")
keys = ["stall_barrier", "stall_long_sb", "stall_lg", "stall_short_sb", "stall_mio", "stall_math", "stall_wait", "stall_membar", "stall_branch_resolving", "stall_not_selected"]
print("# file:line  samples  share  inst_share  top stall reasons | source")
for f, ln, d in sorted(lines, key=lambda x: -int(x[2]["# Samples"]))[:top]:
    s = int(d["# Samples"]);
    rs = sorted(((int(d.get(k, 0) or 0), k[6:]) for k in keys), reverse=True)[:3]
    print(f"{f}:{ln:<5d} {s:7d} {s / tot:6.3f} {int(d['Instructions Executed']) / toti:6.3f}  " + " ".join(f"{k}={v}" for v, k in rs if v) + " | " + d["Source"].strip()[:110])
