#!/bin/bash
# ncu evidence: launch list of the bench command + one full capture of each main kernel (256 streams to keep replays short)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
   python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:audio_kernel|channelize16|cascade_kernel|dc_local' \
   --launch-skip 5 --launch-count 5 -o gpurun_out/prof_main -f python tools/quick_bench.py --streams 256 --steps 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
