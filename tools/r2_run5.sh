#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python tools/time_spec_ctx.py > gpurun_out/r2_ctx5.log 2>&1
ASRK_SPEC_ZSCORE=separate timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err
for w in c3 c4 c5; do
  timeout 900 python bench.py --steps 10 --warmup 3 --workload $w > gpurun_out/r2_bench5_$w.json 2> gpurun_out/r2_bench5_$w.err
done
timeout 900 python bench.py --steps 20 --warmup 5 --surface keras --no-cpu-baseline > gpurun_out/r2_bench5_keras.json 2> gpurun_out/r2_bench5_keras.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench5_ref.json 2> gpurun_out/r2_bench5_ref.err
echo done
