#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
rm -f gpurun_out/parity.json
timeout 2700 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t11.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_t11.log
for w in c3; do
  timeout 900 python bench.py --steps 20 --warmup 5 --workload $w --no-cpu-baseline > gpurun_out/r2_bench11_$w.json 2> gpurun_out/r2_bench11_$w.err
done
echo done
