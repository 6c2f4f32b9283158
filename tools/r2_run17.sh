#!/bin/bash
# round 2, run 17: concurrent alpha/beta lattice CTAs + single-write grad rows (tests, C3 bench, lse variants), PCIe ceilings
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ctc.py tests/test_ctc_first_principles.py tests/test_gpu_parity_configs.py tests/test_properties.py -x -q -m gpu > gpurun_out/r2_t17.log 2>&1
tail -5 gpurun_out/r2_t17.log
timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_b17_c3.log 2>&1; tail -1 gpurun_out/r2_b17_c3.log | cut -c1-600
ASRK_LIB_SUFFIX=_v1 timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 > gpurun_out/r2_b17_c3_v1.log 2>&1; tail -1 gpurun_out/r2_b17_c3_v1.log | cut -c1-600
timeout 600 python tools/pcie_ceiling.py > gpurun_out/r2_pcie17.log 2>&1; cat gpurun_out/r2_pcie17.log
