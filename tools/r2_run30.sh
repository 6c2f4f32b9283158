#!/bin/bash
# merged tail as default: full GPU suite, C2 / C5 / keras benches (+ two-kernel A/B), ncu evidence of the merged step
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t30.log 2>&1
tail -3 gpurun_out/r2_t30.log
for fl in "" "--two-kernel-tail" "--workload c5" "--surface keras"; do
  n=$(echo "$fl" | tr -d ' -')
  timeout 600 python bench.py --steps 30 --warmup 5 $fl > gpurun_out/r2_b30_$n.log 2>&1
  grep '^{' gpurun_out/r2_b30_$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$n', d['ms_per_step'], d['value'], d['kernel_ms'], d['roofline'] and (d['roofline']['frac'], d['roofline']['step_frac']), d['e2e'] and d['e2e']['value'], d['gpu_launches'])"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_c2.csv python tools/profile_step.py 3 c2 > gpurun_out/r2_ncu30_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_c2keras.csv python tools/profile_step.py 3 c2 keras > gpurun_out/r2_ncu30_keras.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'spectrogram_kernel|stats_kernel|normalize_kernel|fused_small_kernel' -s 6 -c 3 -f -o gpurun_out/prof_step_r2 python tools/profile_step.py 3 c2 > gpurun_out/r2_ncu30_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'fused_small_kernel' -s 1 -c 1 -f -o gpurun_out/prof_keras_r2 python tools/profile_step.py 2 c2 keras >> gpurun_out/r2_ncu30_full.log 2>&1
echo done
