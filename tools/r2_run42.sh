#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_roundtrip.py -x -q -m gpu -k steps_in_flight > gpurun_out/r2_t42.log 2>&1
grep -v "^frame #" gpurun_out/r2_t42.log | tail -60 | cut -c1-300
