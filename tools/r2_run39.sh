#!/bin/bash
# several steps in flight (tools/time_overlap.py)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
ASRK_OVERLAP_SWEEP=${SWEEP:-1} timeout 300 python tools/time_overlap.py 240 > gpurun_out/r2_overlap${TAG:-39}.log 2>&1
tail -45 gpurun_out/r2_overlap${TAG:-39}.log
