#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2g_bench_${N}gpu.json 2> gpurun_out/r2g_bench_${N}gpu.err; echo "rc $?"
tail -c 2500 gpurun_out/r2g_bench_${N}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --workload c5 > gpurun_out/r2g_bench_c5_${N}gpu.json 2> gpurun_out/r2g_bench_c5_${N}gpu.err; echo "rc $?"
tail -c 1200 gpurun_out/r2g_bench_c5_${N}gpu.json
nvidia-smi topo -m > gpurun_out/r2g_topo.txt 2>&1; lscpu | head -20 >> gpurun_out/r2g_topo.txt
