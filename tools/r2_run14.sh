#!/bin/bash
# 2 GPUs: on-device multi-rank parity + scaling points
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_mgpu.log 2>&1
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu >> gpurun_out/r2_mgpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_mgpu.log
for n in 1 2; do
  if [ $n = 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 5 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err
  fi
done
echo done
