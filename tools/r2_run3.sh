#!/bin/bash
# round 2, GPU call 3: team-distributed in-kernel z-score + shared-memory diet variants
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_features.py tests/test_gpu_noise.py tests/test_gpu_canaries.py tests/test_loader_class.py tests/test_gpu_logfbank.py -x -q -m gpu > gpurun_out/r2_t3.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_t3.log
ASRK_LIB_SUFFIX=_twp timeout 1200 python -m pytest tests/test_gpu_features.py -x -q -m gpu > gpurun_out/r2_t3b.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_t3b.log
: > gpurun_out/r2_spec3.log
for suf in "" _tw _twp; do
for v in "3 kernel" "3 separate" "2 kernel"; do
  set -- $v
  echo "== lib$suf teams=$1 zscore=$2" >> gpurun_out/r2_spec3.log
  ASRK_LIB_SUFFIX=$suf ASRK_SPEC_TEAMS=$1 ASRK_SPEC_ZSCORE=$2 ASRK_TIME_MODES=fbank,fbank_raw timeout 300 python tools/time_spec.py >> gpurun_out/r2_spec3.log 2>&1
done
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
echo done
