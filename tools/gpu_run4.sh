#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x "$@" 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
