#!/usr/bin/env python
"""Turns ncu artefacts brought back in gpurun_out/ into the small text summaries kept
under profiles/ (the .ncu-rep files themselves are scratch).

  python tools/ncu_summary.py launches <launches.csv> <out.txt>
  python tools/ncu_summary.py rep <file.ncu-rep> <out.txt>
"""
import csv
import io
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_read\.sum|dram__bytes_write\.sum|dram__bytes_read\.sum\.per_second|"
    r"dram__bytes_write\.sum\.per_second|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"lts__t_sector_hit_rate\.pct|l1tex__t_sector_hit_rate\.pct|lts__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__thread_inst_executed_per_inst_executed\.ratio|"
    r"launch__registers_per_thread|launch__grid_size|launch__block_size|launch__shared_mem_per_block.*|"
    r"launch__occupancy_limit_.*|smsp__average_warps_issue_stalled_(long_scoreboard|short_scoreboard|barrier|wait|"
    r"lg_throttle|mio_throttle|math_pipe_throttle|not_selected|branch_resolving|membar)_per_issue_active\.ratio|"
    r"sm__inst_executed_pipe_lsu\.avg\.pct_of_peak_sustained_active|sm__pipe_tensor_cycles_active.*|"
    r"l1tex__data_bank_conflicts_pipe_lsu.*sum|smsp__inst_executed\.sum)$")


def launches(path, out):
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[start]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[start + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        a = agg.setdefault(r[ki].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {path}\n")
        f.write(f"# {'kernel':60s} launches   avg_us   share\n")
        for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{n:62s} {a[0]:8d} {a[1] / a[0]:8.2f} {a[1] / tot:7.3f}\n")
    print(open(out).read())


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on : {path}\n")
        for r in rows[2:]:
            f.write(f"\nkernel: {r[hdr.index('Kernel Name')]}\n")
            for h, u, v in zip(hdr, units, r):
                if KEEP.match(h):
                    f.write(f"  {h:86s} {v:>16s} {u}\n")
    print(open(out).read())


def traffic(out, *reps):
    """profiles/traffic.json: {kernel short name: {dram_bytes_per_launch, source}} from --set full captures."""
    import json
    import os

    table = json.load(open(out)) if os.path.exists(out) else {}
    for path in reps:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        per = {}
        for r in rows[2:]:
            name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").split("<")[0].split("::")[-1]
            tot = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(m)
                tot += float(r[i].replace(",", "")) * scale[units[i]]
            per.setdefault(name, []).append(tot)
        for name, vals in per.items():
            table[name] = {"dram_bytes_per_launch": sum(vals) / len(vals), "launches_captured": len(vals),
                           "source": os.path.basename(path)}
    json.dump(table, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2], *sys.argv[3:])
    else:
        {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2], sys.argv[3])
