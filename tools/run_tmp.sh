nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err; python tools/show_bench.py gpurun_out/bench_n2.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print(d['extra']['icp_sharded'])"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1c.json 2> gpurun_out/bench_n1c.err;  python tools/show_bench.py gpurun_out/bench_n1c.json; python -c "
import json; d=json.load(open('gpurun_out/bench_n1c.json')); print(d['extra']['icp_sharded'])"
