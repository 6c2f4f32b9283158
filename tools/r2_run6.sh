#!/bin/bash
# A/B: round-1 spectrogram kernel (with the round-2 codelets) vs the round-2 rewrite, same contexts
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r2_ab6.log
for suf in _r1 ""; do
  echo "=== lib$suf" >> gpurun_out/r2_ab6.log
  ASRK_LIB_SUFFIX=$suf ASRK_SPEC_ZSCORE=separate ASRK_TIME_MODES=fbank,fbank_raw timeout 300 python tools/time_spec.py >> gpurun_out/r2_ab6.log 2>&1
  ASRK_LIB_SUFFIX=$suf ASRK_SPEC_ZSCORE=separate timeout 300 python tools/time_spec_ctx.py >> gpurun_out/r2_ab6.log 2>&1
  for k in 1 2; do
  ASRK_LIB_SUFFIX=$suf ASRK_SPEC_ZSCORE=separate timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step %.4f'%d['ms_per_step'], {k:round(v,4) for k,v in d['kernel_ms'].items()})" >> gpurun_out/r2_ab6.log 2>&1
  done
done
ASRK_LIB_SUFFIX=_r1 timeout 900 python -m pytest tests/test_gpu_features.py -x -q -m gpu > gpurun_out/r2_t6.log 2>&1
ASRK_LIB_SUFFIX=_r1 ASRK_TIME_MODES=fbank_raw timeout 600 ncu --set full --import-source on --clock-control none -k regex:spectrogram_kernel -s 4 -c 1 -f -o gpurun_out/prof_spec_r2c_r1 python tools/time_spec.py > gpurun_out/r2_ncu6.log 2>&1
ASRK_TIME_MODES=fbank_raw timeout 600 ncu --set full --import-source on --clock-control none -k regex:spectrogram_kernel -s 4 -c 1 -f -o gpurun_out/prof_spec_r2c_new python tools/time_spec.py >> gpurun_out/r2_ncu6.log 2>&1
echo done
