python tools/range_bench.py
