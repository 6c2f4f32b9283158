#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roundtrip.py -x -q -m gpu > gpurun_out/r2_t20.log 2>&1
tail -5 gpurun_out/r2_t20.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_b20_c2.log 2>&1; grep '^{' gpurun_out/r2_b20_c2.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['e2e'])"
