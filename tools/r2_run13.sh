#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
: > gpurun_out/r2_t13.log
for flag in "" "--no-graph"; do
for k in 1 2; do
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline $flag 2>>gpurun_out/r2_t13.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench $flag ms/step %.4f'%d['ms_per_step'], 'launches',d['gpu_launches'], 'e2e',d['e2e']['value'] if d['e2e'] else None)" >> gpurun_out/r2_t13.log 2>&1
done
done
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --workload c5 2>>gpurun_out/r2_t13.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench c5 ms/step %.4f'%d['ms_per_step'], 'launches',d['gpu_launches'], d.get('label_error_mean'))" >> gpurun_out/r2_t13.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c3 2>>gpurun_out/r2_t13.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench c3 ms/step %.4f'%d['ms_per_step'], 'launches',d['gpu_launches'])" >> gpurun_out/r2_t13.log 2>&1
echo done
