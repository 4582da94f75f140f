#!/bin/bash
for cfg in "PMR446_FF_CVT=arith" "PMR446_FF_CVT=lut" "PMR446_FF_CVT=lut PMR446_FF_THREADS=64"; do
  env $cfg python tools/quick_bench.py --streams 1024 --steps 3 2>&1 | tail -1 | sed "s/^/$cfg /"
done
PMR446_FF_CVT=lut python -m pytest tests/test_gpu_pmr_parity.py -m gpu -x -q -k "cfg_b or awkward or large_chunk" 2>&1 | tail -2
