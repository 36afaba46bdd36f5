# Round-2 closing profiling pass (run under gpurun, one GPU). Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py --no-extra --steps 3 --warmup 3 > /dev/null 2>&1   # warm page cache + scan cache
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_bench_r2b.csv python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/ncu_launches_r2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused -s 4 -c 1 -o gpurun_out/prof_vg_fused_r2b -f python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/ncu1_r2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"minmax_bulk|key_kernel|hist_kernel|base_kernel|scatter_kernel|head_count|scan_counts|^reduce_kernel" -s 13 -c 13 -o gpurun_out/prof_vg50m_r2b -f python tools/vg_one.py 1 10 5 > gpurun_out/ncu4_r2b.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 26 -c 26 --csv --log-file gpurun_out/launches_vg50m_r2b.csv python tools/vg_one.py 1 10 5 > gpurun_out/ncu5_r2b.log 2>&1
ls -la gpurun_out | tail -8
