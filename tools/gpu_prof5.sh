#!/bin/bash
# full ncu capture of one launch of each hot kernel (256 streams) + launch list of a short bench run
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k 'regex:cascade_kernel|hbarb_tile|channelize16|audio_fft' \
   --launch-skip 8 --launch-count 4 -o gpurun_out/prof_$1 -f python tools/quick_bench.py --streams 256 --steps 2 > gpurun_out/ncu_full_$1.log 2>&1
tail -2 gpurun_out/ncu_full_$1.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$1.csv \
   python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu_$1.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu_$1.log
