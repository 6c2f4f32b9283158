#!/bin/bash
# full GPU suite, the four workload benches, refreshed ncu evidence for the kernels changed since run 16 (C3, C4)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t24.log 2>&1
tail -3 gpurun_out/r2_t24.log
for w in c2 c3 c4 c5; do
  timeout 600 python bench.py --workload $w --steps 30 --warmup 5 > gpurun_out/r2_b24_$w.log 2>&1
  grep '^{' gpurun_out/r2_b24_$w.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['config']['workload'][:3], d['ms_per_step'], d['value'], d['kernel_ms'], d['roofline'] and (d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['step_frac']), d['e2e'] and d['e2e']['value'], d['gpu_launches'])"
done
timeout 600 python bench.py --surface keras --steps 30 --warmup 5 > gpurun_out/r2_b24_keras.log 2>&1
grep '^{' gpurun_out/r2_b24_keras.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('keras', d['ms_per_step'], d['value'])"
for w in c3 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_$w.csv python tools/profile_step.py 3 $w > gpurun_out/r2_ncu24_$w.log 2>&1
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'rows_kernel|lattice_kernel|grad_kernel' -s 3 -c 3 -f -o gpurun_out/prof_c3_r2 python tools/profile_step.py 2 c3 > gpurun_out/r2_ncu24_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'snr2k_kernel|spectrogram_kernel' -s 2 -c 2 -f -o gpurun_out/prof_c4_r2 python tools/profile_step.py 2 c4 >> gpurun_out/r2_ncu24_full.log 2>&1
echo done
