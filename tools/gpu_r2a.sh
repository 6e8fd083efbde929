#!/bin/bash
# round-2 first GPU pass: parity tests, driver-style bench line, mujoco probe, steady-state DRAM traffic (1-pass ncu, caches untouched)
mkdir -p gpurun_out
python -c "import mujoco" > gpurun_out/r2a_mujoco_probe.log 2>&1; python -c "import dm_control" >> gpurun_out/r2a_mujoco_probe.log 2>&1
nproc >> gpurun_out/r2a_mujoco_probe.log; nvidia-smi -L >> gpurun_out/r2a_mujoco_probe.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc $?"
tail -c 3000 gpurun_out/r2a_bench.json
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2a_bench_ref.json 2>&1
# steady-state DRAM traffic of the step kernel: single-pass metrics, no cache flush between launches (--cache-control none)
CMD="python bench.py --steps 20 --warmup 3 --preroll 300 --no-cpu-baseline --no-extras --no-graph"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:step_kernel -s 2440 -c 80 --csv --log-file gpurun_out/r2a_dram_steady.csv $CMD > gpurun_out/r2a_dram_steady.log 2>&1
tail -3 gpurun_out/r2a_dram_steady.csv
