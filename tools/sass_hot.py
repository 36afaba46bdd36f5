#!/usr/bin/env python
"""Hot SASS lines of an ncu report: python tools/sass_hot.py <rep> [min_share]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]].replace(",", ""))
    except Exception: return 0.0
ti = sum(f(r, "Instructions Executed") for r in data); ts = sum(f(r, "# Samples") for r in data)
tt = sum(f(r, "Thread Instructions Executed") for r in data)
print(f"warp inst {ti/1e9:.3f} G   thread inst {tt/1e9:.3f} G   avg lanes {tt/ti:.2f}   samples {ts:.0f}")
for n, r in enumerate(data):
    inst = f(r, "Instructions Executed"); s = f(r, "# Samples")
    if inst > ti * thr or s > ts * thr:
        print(f"{n:4d} {r[ix['Source']][:64]:64s} inst {inst/1e6:8.1f}M lanes {r[ix['Avg. Threads Executed']]:>5s} samp {s/ts*100:5.1f}% longsb {f(r,'stall_long_sb')/ts*100:5.1f}%")
