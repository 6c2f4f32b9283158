#!/bin/bash
# round 2 evidence: launch lists and --set full captures of the step's kernels (C2, C2 keras surface, C3, C4)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for w in c2 c3 c4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_$w.csv python tools/profile_step.py 3 $w > gpurun_out/r2_ncu16_$w.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_c2keras.csv python tools/profile_step.py 3 c2 keras > gpurun_out/r2_ncu16_keras.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'spectrogram_kernel|stats_kernel|normalize_kernel|fused_small_kernel' -s 8 -c 4 -f -o gpurun_out/prof_step_r2 python tools/profile_step.py 3 c2 > gpurun_out/r2_ncu16_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'rows_kernel|lattice_kernel|grad_kernel' -s 3 -c 3 -f -o gpurun_out/prof_c3_r2 python tools/profile_step.py 2 c3 >> gpurun_out/r2_ncu16_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'snr2k_kernel|spectrogram_kernel' -s 2 -c 2 -f -o gpurun_out/prof_c4_r2 python tools/profile_step.py 2 c4 >> gpurun_out/r2_ncu16_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'fused_small_kernel' -s 1 -c 1 -f -o gpurun_out/prof_keras_r2 python tools/profile_step.py 2 c2 keras >> gpurun_out/r2_ncu16_full.log 2>&1
echo done
