#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roundtrip.py tests/test_gpu_ctc.py tests/test_gpu_features.py -x -q -m gpu > gpurun_out/r2_t25.log 2>&1
tail -5 gpurun_out/r2_t25.log
for fl in "" "--two-kernel-tail"; do
timeout 600 python bench.py --steps 30 --warmup 5 $fl > gpurun_out/r2_b25_c2$fl.log 2>&1; grep '^{' gpurun_out/r2_b25_c2$fl.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['config']['tail'][:12], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d['gpu_launches'], d['loss_mean'])"
done
timeout 600 python bench.py --steps 30 --warmup 5 --surface keras > gpurun_out/r2_b25_keras.log 2>&1; grep '^{' gpurun_out/r2_b25_keras.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('keras', d['ms_per_step'], d['loss_mean'])"
timeout 600 python bench.py --steps 30 --warmup 5 --workload c5 > gpurun_out/r2_b25_c5.log 2>&1; grep '^{' gpurun_out/r2_b25_c5.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('c5', d['ms_per_step'], d['loss_mean'])"
