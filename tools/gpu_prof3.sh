#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:hb_arb|channelize16' \
   --launch-skip 2 --launch-count 2 -o gpurun_out/prof_c -f python tools/quick_bench.py --streams 256 --steps 1 > gpurun_out/ncu_full_c.log 2>&1
tail -2 gpurun_out/ncu_full_c.log
