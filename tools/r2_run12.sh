#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_features.py -x -q -m gpu > gpurun_out/r2_t12.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_t12.log
for k in 1 2; do
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench ms/step %.4f'%d['ms_per_step'], {k:round(v,4) for k,v in d['kernel_ms'].items()})" >> gpurun_out/r2_t12.log 2>&1
done
ASRK_TIME_MODES=fbank,fbank_raw timeout 300 python tools/time_spec.py >> gpurun_out/r2_t12.log 2>&1
echo done
