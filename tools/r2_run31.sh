#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for fl in "--no-graph" "--no-graph --workload c5" ""; do
  n=$(echo "$fl" | tr -d ' -')
  timeout 600 python bench.py --steps 20 --warmup 5 $fl > gpurun_out/r2_b31_$n.log 2>&1
  grep '^{' gpurun_out/r2_b31_$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$n', d['ms_per_step'], d['kernel_ms']['ctc_fused'], d['e2e'] and (d['e2e']['value'], d['e2e']['results_stay_on_device']['value']), d['gpu_launches'])"
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
