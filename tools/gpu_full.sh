#!/bin/bash
# full round evidence: gpu tests, smoke, bench (both arms), launch list, full ncu capture
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_ref.json
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_ours.json
cut -c1-1500 gpurun_out/bench_ours.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
   python bench.py --steps 2 --warmup 1 > gpurun_out/bench_under_ncu.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:audio_kernel|channelize16|cascade_kernel' \
   --launch-skip 4 --launch-count 4 -o gpurun_out/prof_main -f python tools/quick_bench.py --streams 256 --steps 1 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
