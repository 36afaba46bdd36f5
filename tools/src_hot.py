#!/usr/bin/env python
"""Hot CUDA source lines of an ncu report (needs -lineinfo): python tools/src_hot.py <rep> [min_share]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur_file = ""; agg = {}; hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr) or r[0] == "": continue   # sass rows have empty Line No
    try:
        s = float(r[ix["# Samples"]].replace(",", "")); inst = float(r[ix["Instructions Executed"]].replace(",", ""))
    except ValueError: continue
    agg[(cur_file, r[0], r[1])] = (s, inst)
ts = sum(v[0] for v in agg.values()); ti = sum(v[1] for v in agg.values())
print(f"samples {ts:.0f}  warp-inst {ti/1e6:.1f}M")
for (f, ln, src), (s, inst) in sorted(agg.items(), key=lambda x: -x[1][0]):
    if s > ts * thr: print(f"{s/ts*100:5.1f}%  inst {inst/ti*100:5.1f}%  {f}:{ln:>4s}  {src.strip()[:100]}")
