#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x -m gpu \
  "tests/test_gpu_pmr_parity.py::test_tiny_chunks_and_empty_call" "tests/test_gpu_pmr_parity.py::test_2400k_awkward_chunks_exercise_tile_edges" \
  "tests/test_gpu_pmr_parity.py::test_waterfall_rows_match" "tests/test_gpu_pmr_parity.py::test_selector_taps_rssi_and_edge_samples" \
  "tests/test_gpu_round2.py::test_waterfall_width_with_large_prime_factor" "tests/test_gpu_round2.py::test_other_resampler_plans" \
  "tests/test_gpu_dsd_parity.py" "tests/test_gpu_receiver_parity.py" > gpurun_out/sanitizer_memcheck_r02t.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck_r02t.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -x -m gpu \
  "tests/test_gpu_pmr_parity.py::test_2400k_awkward_chunks_exercise_tile_edges" "tests/test_gpu_pmr_parity.py::test_waterfall_rows_match" \
  "tests/test_gpu_dsd_parity.py" > gpurun_out/sanitizer_racecheck_r02t.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck_r02t.log
tail -4 gpurun_out/sanitizer_memcheck_r02t.log; tail -6 gpurun_out/sanitizer_racecheck_r02t.log
