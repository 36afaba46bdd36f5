"""Extracts the PCD golden vectors of the reference's own test (pc/io_test.go:16-216) into
tests/golden/pcd_cases.json (byte strings hex-encoded).  Run in the container that has
/root/reference; the GPU box only reads the committed JSON."""
import json
import os
import re

SRC = "/root/reference/pc/io_test.go"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pcd_cases.json")

text = open(SRC).read()
body = text[text.index("testCases := map[string]struct {"):text.index("for name, tt := range testCases")]
cases = {}
# name: { pcd: []byte(`...`) | []byte{...} | []byte("..."), [expected: ...,] [err: X,] }
for m in re.finditer(r'"(\w+)": \{\s*pcd: \[\]byte(\(`(.*?)`\)|\{(.*?)\}|\("(.*?)"\)),(.*?)\n\t\t\},', body, re.S):
    name, _, raw, hexes, quoted, rest = m.groups()
    if raw is not None:
        data = raw.encode()
    elif hexes is not None:
        data = bytes(int(h, 16) for h in re.findall(r"0x([0-9a-fA-F]{2})", hexes))
    else:
        data = quoted.encode()
    err = re.search(r"err:\s*([\w.]+)", rest)
    pts = [[float(a), float(b), float(c), int(d)] for a, b, c, d in
           re.findall(r"\{(-?[\d.]+), (-?[\d.]+), (-?[\d.]+), (\d+)\}", rest)]
    cases[name] = {"pcd_hex": data.hex(), "expected": pts, "err": err.group(1) if err else None}
assert {"Ascii", "Binary", "BinaryCompressed", "ErrorVersion", "ErrorBinaryCompressedInvalidData"} <= set(cases), cases.keys()
json.dump({"source": "pc/io_test.go:16-216", "cases": cases}, open(OUT, "w"), indent=1)
print(len(cases), "cases ->", OUT)
