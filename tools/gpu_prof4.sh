#!/bin/bash
# usage: gpu_prof4.sh <kernel-regex> <tag>
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$1" \
   --launch-skip 2 --launch-count 1 -o gpurun_out/prof_$2 -f python tools/quick_bench.py --streams 256 --steps 1 > gpurun_out/ncu_full_$2.log 2>&1
tail -2 gpurun_out/ncu_full_$2.log
