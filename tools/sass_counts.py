#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libpcgol_b200.so (cuobjdump -sass): the bulk-async / barrier / matching
instructions that show which kernels use the sm_100a data-movement features.
    python tools/sass_counts.py > profiles/<round>_sass_counts.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "pcgol_b200/libpcgol_b200.so"
WANT = ["UBLKCP", "SYNCS", "UTMALDG", "LDGSTS", "MATCH", "REDUX", "ATOMS", "ATOMG", "RED", "BAR", "LDG", "STG", "LDS", "STS",
        "SHFL", "VOTE", "FMNMX", "MUFU"]
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
name = None
counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        counts[name]["_total"] += 1
        for w in WANT:
            if op == w or op.startswith(w + "."):
                counts[name][w] += 1
print(f"# cuobjdump -sass {LIB}: instructions per kernel (static counts)")
print(f"# {'kernel':58s} {'total':>6s} " + " ".join(f"{w:>7s}" for w in WANT))
for k, c in counts.items():
    print(f"{k[-60:]:60s} {c['_total']:6d} " + " ".join(f"{c[w]:7d}" for w in WANT))
