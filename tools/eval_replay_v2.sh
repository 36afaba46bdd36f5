#!/bin/bash
# Evaluate the strict-replay variant (-DPCG_REPLAY_V2) against the default build.
# Here (CPU):   bash tools/eval_replay_v2.sh build
# On the GPU:   gpurun --timeout 300 -- 'bash tools/eval_replay_v2.sh run'
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
cd "$root"
case "$1" in
  build) bash tools/build_variant.sh v2 -DPCG_REPLAY_V2 ;;
  run)
    mkdir -p gpurun_out
    lib=$root/build_variants/libpcg_v2.so
    echo "== parity (variant)"; PCG_LIB=$lib python -m pytest tests/test_gpu_edges.py tests/test_gpu_parity.py tests/test_gpu_next.py -x -q -k "strict or icp or replay" 2>&1 | tail -2
    echo "== walk cycles per stream: default, variant, variant with the old walk"
    python tools/replay_stats.py 2>&1 | tail -6
    PCG_LIB=$lib python tools/replay_stats.py 2>&1 | tail -6
    PCG_LIB=$lib PCG_REPLAY_V1=1 python tools/replay_stats.py 2>&1 | tail -6
    echo "== Fit"
    python bench.py --only icp 2>/dev/null > gpurun_out/icp_default.json
    PCG_LIB=$lib python bench.py --only icp 2>/dev/null > gpurun_out/icp_v2.json
    python - <<'PY'
import json
for v in ("default", "v2"):
    s = json.load(open(f"gpurun_out/icp_{v}.json"))["modes"]["strict"]
    print(v, round(s["value"], 1), "alignments/s", {k: round(x["avg_us"], 1) for k, x in s["kernels"].items()})
PY
    ;;
  *) echo "usage: $0 build|run"; exit 1 ;;
esac
