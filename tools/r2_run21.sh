#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roundtrip.py tests/test_gpu_ctc.py -x -q -m gpu > gpurun_out/r2_t21.log 2>&1
tail -5 gpurun_out/r2_t21.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_b21_c2.log 2>&1; grep '^{' gpurun_out/r2_b21_c2.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['e2e'])"
for sfx in "" _g2; do
ASRK_LIB_SUFFIX=$sfx timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_b21_c3$sfx.log 2>&1; grep '^{' gpurun_out/r2_b21_c3$sfx.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['kernel_ms'])"
done
