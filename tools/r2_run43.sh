#!/bin/bash
# final single-GPU check of the tree as shipped: whole GPU suite, smoke, default bench, reference arm (short), C3 with two steps in flight
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export ASRK_BENCH_CACHE=/tmp/asrk_cache
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t43.log 2>&1
tail -2 gpurun_out/r2_t43.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/r2_b43_n1.log 2>&1
grep '^{' gpurun_out/r2_b43_n1.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['steps'], d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['step_frac'], d['e2e']['value'], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'])"
timeout 400 python bench.py --workload c3 --steps-in-flight 2 --steps 12 --no-cpu-baseline > gpurun_out/r2_b43_c3_f2.log 2>&1
grep '^{' gpurun_out/r2_b43_c3_f2.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('c3 in flight', d['config']['steps_in_flight'], d['ms_per_step'], d['value'], d['kernel_ms'])"
tail -2 gpurun_out/r2_b43_c3_f2.log | cut -c1-300
