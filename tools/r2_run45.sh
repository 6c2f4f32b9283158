#!/bin/bash
# default bench (100 steps, two steps in flight), CPU baseline skipped to stay inside the last GPU minute
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 60 python bench.py --no-cpu-baseline > gpurun_out/r2_b45_n1.log 2>&1
grep '^{' gpurun_out/r2_b45_n1.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['steps'], d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['step_frac'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
tail -2 gpurun_out/r2_b45_n1.log | cut -c1-200
