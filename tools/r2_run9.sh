#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests/test_gpu_ctc.py tests/test_ctc_first_principles.py tests/test_gpu_parity_configs.py tests/test_gpu_canaries.py -x -q -m gpu > gpurun_out/r2_t9.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_t9.log
for w in c2 c3; do
  timeout 900 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/r2_bench9_$w.json 2> gpurun_out/r2_bench9_$w.err
done
echo done
