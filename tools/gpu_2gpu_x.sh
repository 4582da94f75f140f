#!/bin/bash
# 2-GPU check on the final kernels: NCCL shard-equality test + weak-scaling bench line
mkdir -p gpurun_out
nvidia-smi -L | head -4
python -m pytest tests/test_gpu_shard_nccl.py -m gpu -x -q > gpurun_out/pytest_nccl_2gpu_r02x.log 2>&1; echo "exit $?" >> gpurun_out/pytest_nccl_2gpu_r02x.log; tail -4 gpurun_out/pytest_nccl_2gpu_r02x.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02x_2gpu.json 2> gpurun_out/bench_r02x_2gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02x_2gpu.json'))
print('2gpu', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
PY
