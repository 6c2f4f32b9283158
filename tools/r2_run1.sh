#!/bin/bash
# round 2, GPU call 1: parity of the rewritten spectrogram kernel + timing of its variants
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_smi.log 2>&1
timeout 1200 python -m pytest tests/test_gpu_features.py tests/test_gpu_noise.py tests/test_gpu_canaries.py -x -q -m gpu > gpurun_out/r2_t1.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_t1.log
: > gpurun_out/r2_spec.log
for v in "3 kernel" "3 separate" "2 kernel" "2 separate"; do
  set -- $v
  echo "== conv1 teams=$1 zscore=$2" >> gpurun_out/r2_spec.log
  ASRK_SPEC_TEAMS=$1 ASRK_SPEC_ZSCORE=$2 timeout 300 python tools/time_spec.py >> gpurun_out/r2_spec.log 2>&1
done
echo "== conv0 teams=3 zscore=kernel" >> gpurun_out/r2_spec.log
ASRK_LIB_SUFFIX=_c0 timeout 300 python tools/time_spec.py >> gpurun_out/r2_spec.log 2>&1
echo "== conv0 teams=2 zscore=kernel" >> gpurun_out/r2_spec.log
ASRK_LIB_SUFFIX=_c0 ASRK_SPEC_TEAMS=2 timeout 300 python tools/time_spec.py >> gpurun_out/r2_spec.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
echo done
