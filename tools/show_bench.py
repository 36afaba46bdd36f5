#!/usr/bin/env python
"""Compact view of a bench.py JSON line."""
import json
import sys

d = json.load(open(sys.argv[1]))
ex = d.get("extra", {})
print(f"{d['metric']}: value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:.1f} "
      f"({d['e2e'].get('ms_per_step', 0):.4f} ms)  launches {d.get('gpu_launches')}  cpu {d.get('cpu_baseline', {}).get('value')}")
print("  clocks", d.get("clocks"))
r = d.get("roofline") or {}
print(f"  roofline: {r.get('kernel')} avg {r.get('avg_launch_us')} us achieved {r.get('achieved')} frac {r.get('frac')}")
print("  pipeline:", d.get("pipeline_roofline"))
for k, v in (d.get("kernels") or {}).items():
    print(f"    {k:44s} {v['launches']:4d} x {v['avg_us']:9.2f} us  share {v['share']:.3f}")
nn = ex.get("nn")
if nn:
    print(f"NN: {nn['value'] / 1e6:.1f} Mq/s ({nn['ms_per_step']:.3f} ms/10M)  e2e {nn['e2e']['value'] / 1e6:.1f} Mq/s  build {nn['index_build_ms']}"
          f"  cpu1 {nn.get('cpu_baseline', {}).get('value')}  parity {nn.get('parity_sample')}")
    for k, v in (nn.get("kernels") or {}).items():
        print(f"    {k:44s} {v['launches']:4d} x {v['avg_us']:9.2f} us  share {v['share']:.3f}")
icp = ex.get("icp")
if icp:
    for m, v in icp["modes"].items():
        print(f"ICP {m}: {v['value']:.1f} align/s ({v['ms_per_alignment']:.3f} ms, {v['iterations']} it)  e2e {v['e2e']['value']:.1f}")
        for k, kv in (v.get("kernels") or {}).items():
            print(f"    {k:44s} {kv['launches']:4d} x {kv['avg_us']:9.2f} us  share {kv['share']:.3f}")
    print("  cpu", icp.get("cpu_baseline", {}).get("value"), "parity", icp.get("parity"))

for k in ("icp_sharded", "icp_farm"):
    if ex.get(k):
        v = dict(ex[k]); v.pop("trans", None); v.pop("config", None)
        print(k, v)
