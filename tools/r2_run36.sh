#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_noise.py tests/test_gpu_features.py tests/test_gpu_parity_configs.py -x -q -m gpu 2>&1 | tail -1
timeout 600 python bench.py --workload c4 --steps 20 --warmup 5 > gpurun_out/r2_b36_c4.log 2>&1
grep '^{' gpurun_out/r2_b36_c4.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['ms_per_step'], d['value'], d['kernel_ms'])"
