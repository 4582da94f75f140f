#!/bin/bash
# round 2, GPU call A: FFMA2 issue-rate probe, the whole GPU parity suite (old + new cases), compute-sanitizer on a subset, bench sanity
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_r02a.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ffma2_probe tools/ffma2_probe.cu && /tmp/ffma2_probe > gpurun_out/ffma2_probe_r02a.jsonl 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02a.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_r02a.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x -m gpu \
  "tests/test_gpu_pmr_parity.py::test_tiny_chunks_and_empty_call" "tests/test_gpu_pmr_parity.py::test_2400k_awkward_chunks_exercise_tile_edges" \
  "tests/test_gpu_round2.py::test_waterfall_large_widths" "tests/test_gpu_round2.py::test_waterfall_width_with_large_prime_factor" \
  "tests/test_gpu_dsd_parity.py" > gpurun_out/sanitizer_memcheck_r02a.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck_r02a.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
tail -3 gpurun_out/pytest_gpu_r02a.log; tail -3 gpurun_out/sanitizer_memcheck_r02a.log; cat gpurun_out/ffma2_probe_r02a.jsonl; cut -c1-600 gpurun_out/bench_r02a.json
