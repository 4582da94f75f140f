#!/bin/bash
# final round-2 evidence, part B: ncu full captures of the kernels changed since r02v + sanitizer on the tests that exercise them
T=${1:-r02w}
mkdir -p gpurun_out
bash tools/gpu_prof6.sh audio_fft4 af4_$T python tools/quick_bench.py --streams 512 --steps 1
DSD_STREAMS=256 bash tools/gpu_prof6.sh dsd_backend dsb_$T python tools/probe_dsd.py
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest -q -x -m gpu \
  "tests/test_gpu_pmr_parity.py::test_tiny_chunks_and_empty_call" "tests/test_gpu_pmr_parity.py::test_2400k_awkward_chunks_exercise_tile_edges" \
  "tests/test_gpu_pmr_parity.py::test_five_channels_odd_row_count" "tests/test_gpu_pmr_parity.py::test_fir_deemphasis_variant" \
  "tests/test_gpu_dsd_parity.py" > gpurun_out/sanitizer_memcheck_$T.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck_$T.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest -q -x -m gpu \
  "tests/test_gpu_pmr_parity.py::test_2400k_awkward_chunks_exercise_tile_edges" "tests/test_gpu_pmr_parity.py::test_five_channels_odd_row_count" \
  "tests/test_gpu_dsd_parity.py" > gpurun_out/sanitizer_racecheck_$T.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck_$T.log
tail -4 gpurun_out/sanitizer_memcheck_$T.log; tail -5 gpurun_out/sanitizer_racecheck_$T.log
