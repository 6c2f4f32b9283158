#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_ctc.py tests/test_ctc_first_principles.py tests/test_gpu_parity_configs.py tests/test_gpu_canaries.py -x -q -m gpu > gpurun_out/r2_t8.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_t8.log
ASRK_LIB_SUFFIX=_tim timeout 300 python tools/ctc_phase_times.py > gpurun_out/r2_ctcphase8.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:lattice_kernel -s 2 -c 1 -f -o gpurun_out/prof_lattice_r2a \
   python bench.py --steps 3 --warmup 3 --workload c3 --no-cpu-baseline > gpurun_out/r2_ncu8.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:grad_kernel -s 2 -c 1 -f -o gpurun_out/prof_grad_r2a \
   python bench.py --steps 3 --warmup 3 --workload c3 --no-cpu-baseline >> gpurun_out/r2_ncu8.log 2>&1
echo done
