#!/bin/bash
# ncu evidence of the C2 step as shipped (two-kernel tail) and with --merged-tail
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_c2.csv python tools/profile_step.py 3 c2 > gpurun_out/r2_ncu33_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_c2keras.csv python tools/profile_step.py 3 c2 keras > gpurun_out/r2_ncu33_keras.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2_c2merged.csv python tools/profile_step.py 3 c2 logits merged > gpurun_out/r2_ncu33_merged.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'spectrogram_kernel|stats_kernel|normalize_kernel|fused_small_kernel' -s 8 -c 4 -f -o gpurun_out/prof_step_r2 python tools/profile_step.py 3 c2 > gpurun_out/r2_ncu33_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'fused_small_kernel' -s 1 -c 1 -f -o gpurun_out/prof_keras_r2 python tools/profile_step.py 2 c2 keras >> gpurun_out/r2_ncu33_full.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'fused_small_kernel' -s 1 -c 1 -f -o gpurun_out/prof_merged_r2 python tools/profile_step.py 2 c2 logits merged >> gpurun_out/r2_ncu33_full.log 2>&1
echo done
