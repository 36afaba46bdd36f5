# Round-1 (second session) profiling pass: run under gpurun, one GPU. Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
python bench.py --no-extra --steps 3 --warmup 3 > /dev/null 2>&1   # warm page cache + scan cache
ncu --set full --clock-control none --import-source on -k regex:fused -s 4 -c 1 -o gpurun_out/prof_vg_fused_r1b -f python bench.py --no-extra --steps 2 --warmup 3 > gpurun_out/ncu1b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:nearest_simple -s 1 -c 1 -o gpurun_out/prof_nearest_r1b -f python bench.py --only nn --steps 3 --warmup 3 > gpurun_out/ncu2b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:icp_ -s 6 -c 5 -o gpurun_out/prof_icp_r1b -f python bench.py --only icp --steps 3 --warmup 3 > gpurun_out/ncu4b.log 2>&1
ls -la gpurun_out | tail -6
