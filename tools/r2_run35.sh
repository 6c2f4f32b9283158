#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_roundtrip.py -x -q -m gpu 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --merged-tail --steps 30 --warmup 5 > gpurun_out/r2_b35_merged.log 2>&1
grep '^{' gpurun_out/r2_b35_merged.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['config']['tail'][:10], d['ms_per_step'], d['kernel_ms']['ctc_fused'], d['e2e']['value'], d['gpu_launches'])"
