#!/bin/bash
# steps in flight: the new GPU test, then bench.py c2 (default: two steps in flight), c2 serial, c5
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export ASRK_BENCH_CACHE=/tmp/asrk_cache
timeout 300 python -m pytest tests/test_gpu_roundtrip.py -x -q -m gpu 2>&1 | tail -3
show() { grep '^{' $1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['config']['steps_in_flight'], d['config']['feature_ctas'], d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['step_frac'], d['roofline'].get('in_step'), d['kernel_ms'], d['loss_mean'], d['e2e'] and d['e2e']['value'], d['gpu_launches'], d['clocks'], d.get('label_error_mean'))"; }
timeout 400 python bench.py > gpurun_out/r2_b41_c2.log 2>&1; show gpurun_out/r2_b41_c2.log
timeout 300 python bench.py --steps-in-flight 1 --no-cpu-baseline > gpurun_out/r2_b41_c2_serial.log 2>&1; show gpurun_out/r2_b41_c2_serial.log
timeout 300 python bench.py --workload c5 --no-cpu-baseline > gpurun_out/r2_b41_c5.log 2>&1; show gpurun_out/r2_b41_c5.log
tail -3 gpurun_out/r2_b41_c2.log | cut -c1-600
