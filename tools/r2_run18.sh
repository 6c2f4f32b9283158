#!/bin/bash
# round 2, run 18: slim grad staging, snr2k rewrite, staging-kernel grids; tests + C3/C4/C2 bench + PCIe ceilings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_noise.py tests/test_gpu_ctc.py tests/test_ctc_first_principles.py tests/test_gpu_parity_configs.py tests/test_gpu_features.py -x -q -m gpu > gpurun_out/r2_t18.log 2>&1
tail -5 gpurun_out/r2_t18.log
for w in c3 c4; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 > gpurun_out/r2_b18_$w.log 2>&1; grep '^{' gpurun_out/r2_b18_$w.log | cut -c1-700
done
timeout 600 python tools/pcie_ceiling.py > gpurun_out/r2_pcie18.log 2>&1; tail -4 gpurun_out/r2_pcie18.log
ASRK_LIB_SUFFIX=_s4 timeout 600 python tools/pcie_ceiling.py > gpurun_out/r2_pcie18_s4.log 2>&1; tail -4 gpurun_out/r2_pcie18_s4.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2_b18_c2.log 2>&1; grep '^{' gpurun_out/r2_b18_c2.log | cut -c1-2500
