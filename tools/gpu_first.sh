#!/bin/bash
# first GPU contact: parity tests, then a timing probe and a launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py --streams 256 --steps 3 2>&1 | tail -12 | tee gpurun_out/quick256.log
timeout 300 python tools/quick_bench.py --streams 1024 --steps 3 2>&1 | tail -12 | tee gpurun_out/quick1024.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_first.csv python tools/quick_bench.py --streams 256 --steps 1 > gpurun_out/ncu_first.log 2>&1
tail -5 gpurun_out/ncu_first.log
