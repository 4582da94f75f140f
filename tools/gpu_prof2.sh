#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:audio_kernel|channelize16|cascade_kernel' \
   --launch-skip 4 --launch-count 4 -o gpurun_out/prof_b -f python tools/quick_bench.py --streams 256 --steps 1 > gpurun_out/ncu_full_b.log 2>&1
tail -3 gpurun_out/ncu_full_b.log
