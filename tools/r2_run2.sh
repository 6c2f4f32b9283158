#!/bin/bash
# round 2, GPU call 2: ncu --set full of the rewritten spectrogram kernel (raw mode: transform only)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export ASRK_TIME_MODES=fbank_raw
timeout 300 python tools/time_spec.py > gpurun_out/r2_spec2.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:spectrogram_kernel -s 4 -c 1 \
  -f -o gpurun_out/prof_spec_r2a python tools/time_spec.py > gpurun_out/r2_ncu2.log 2>&1
export ASRK_TIME_MODES=fbank
timeout 900 ncu --set full --import-source on --clock-control none -k regex:spectrogram_kernel -s 4 -c 1 \
  -f -o gpurun_out/prof_spec_r2b python tools/time_spec.py >> gpurun_out/r2_ncu2.log 2>&1
echo done
