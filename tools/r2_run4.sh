#!/bin/bash
# round 2, GPU call 4: spectrogram variants (statistics in kernel, separate out tile), new CTC / loader / parity tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.json
: > gpurun_out/r2_spec4.log
for v in "3 kernel 0" "3 separate 0" "3 kernel 1" "2 kernel 0"; do
  set -- $v
  echo "== teams=$1 zscore=$2 sepout=$3" >> gpurun_out/r2_spec4.log
  ASRK_SPEC_TEAMS=$1 ASRK_SPEC_ZSCORE=$2 ASRK_SPEC_SEPOUT=$3 ASRK_TIME_MODES=fbank,fbank_raw timeout 300 python tools/time_spec.py >> gpurun_out/r2_spec4.log 2>&1
done
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t4.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_t4.log
ASRK_SPEC_SEPOUT=1 timeout 600 python -m pytest tests/test_gpu_features.py -x -q -m gpu > gpurun_out/r2_t4b.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_t4b.log
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
ASRK_SPEC_SEPOUT=1 timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench4b.json 2> gpurun_out/r2_bench4b.err
echo done
