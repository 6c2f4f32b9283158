#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roundtrip.py tests/test_gpu_ctc.py -x -q -m gpu 2>&1 | tail -1
timeout 600 python tools/time_zcowork.py 2>&1 | tail -2
timeout 600 python tools/check_merged.py 2>&1 | grep "merged tail"
for fl in "" "--no-graph" "--no-graph --workload c5" "--workload c5"; do
  n=$(echo "$fl" | tr -d ' -')
  timeout 600 python bench.py --steps 20 --warmup 5 $fl > gpurun_out/r2_b32_$n.log 2>&1
  grep '^{' gpurun_out/r2_b32_$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$n', d['ms_per_step'], d['kernel_ms']['ctc_fused'], d['e2e'] and (d['e2e']['value'], d['e2e']['results_stay_on_device']['value']), d['gpu_launches'])"
done
