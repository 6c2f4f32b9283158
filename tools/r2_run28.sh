#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roundtrip.py -x -q -m gpu > gpurun_out/r2_t28.log 2>&1
tail -3 gpurun_out/r2_t28.log
timeout 600 python tools/time_zcowork.py > gpurun_out/r2_zc28.log 2>&1; tail -4 gpurun_out/r2_zc28.log
timeout 600 python tools/check_merged.py > gpurun_out/r2_merged28.log 2>&1; tail -4 gpurun_out/r2_merged28.log
