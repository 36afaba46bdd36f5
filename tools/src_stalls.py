#!/usr/bin/env python
"""Per-source-line stall breakdown of an ncu report: python tools/src_stalls.py <rep> [file-substring] [min_share]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.008
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
cur = ""; ix = None; agg = {}
stalls = ["stall_barrier", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_lg", "stall_mio", "stall_math", "stall_branch_resolving", "stall_not_selected", "stall_selected", "stall_membar", "stall_no_inst", "stall_dispatch"]
for r in rows:
    if len(r) >= 2 and r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No": ix = {h: i for i, h in enumerate(r)}; continue
    if ix is None or len(r) < len(ix) or r[0] == "": continue
    try: s = float(r[ix["# Samples"]].replace(",", ""))
    except ValueError: continue
    d = agg.setdefault((cur, int(r[0]), r[1].strip()[:70]), {"s": 0.0, "inst": 0.0, **{k: 0.0 for k in stalls}, "conf": 0.0})
    d["s"] += s
    try: d["inst"] += float(r[ix["Instructions Executed"]].replace(",", "") or 0)
    except ValueError: pass
    for k in stalls:
        try: d[k] += float(r[ix[k]].replace(",", "") or 0)
        except (ValueError, KeyError): pass
    try: d["conf"] += float(r[ix["L1 Wavefronts Shared Excessive"]].replace(",", "") or 0)
    except ValueError: pass
ts = sum(v["s"] for v in agg.values())
print(f"total samples {ts:.0f}")
for (f, ln, src), v in sorted(agg.items(), key=lambda x: (x[0][0], x[0][1])):
    if want and want not in f: continue
    if v["s"] < ts * thr: continue
    top = sorted(((v[k], k[6:]) for k in stalls), reverse=True)[:3]
    print(f"{v['s'] / ts * 100:5.1f}%  {f}:{ln:4d}  inst {v['inst'] / 1e3:7.0f}k  smem-excess {v['conf'] / 1e3:6.0f}k  " + " ".join(f"{n}={x / max(v['s'], 1) * 100:.0f}%" for x, n in top if x > 0) + f"  | {src}")
