#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for sfx in _zp20 _zp40; do
echo "== $sfx"
ASRK_LIB_SUFFIX=$sfx timeout 600 python -m pytest tests/test_gpu_roundtrip.py -x -q -m gpu 2>&1 | tail -1
ASRK_LIB_SUFFIX=$sfx timeout 600 python tools/time_zcowork.py 2>&1 | tail -1
ASRK_LIB_SUFFIX=$sfx timeout 600 python tools/check_merged.py 2>&1 | grep "merged tail True"
done
