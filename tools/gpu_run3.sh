#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py --streams 1024 --steps 3 2>&1 | tail -8 | tee gpurun_out/quick1024.log
