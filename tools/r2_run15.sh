#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export ASRK_BENCH_WATCHDOG=200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2_scale_n2.json 2> gpurun_out/r2_scale_n2.err
echo "exit $?" >> gpurun_out/r2_scale_n2.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --no-graph > gpurun_out/r2_scale_n2_nograph.json 2> gpurun_out/r2_scale_n2_nograph.err
echo "exit $?" >> gpurun_out/r2_scale_n2_nograph.err
timeout 900 python -m pytest tests/test_gpu_ctc.py tests/test_gpu_parity_configs.py tests/test_properties.py -x -q -m gpu > gpurun_out/r2_t15.log 2>&1
echo "pytest exit $?" >> gpurun_out/r2_t15.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload c3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench c3 ms/step %.4f'%d['ms_per_step'], {k:round(v,4) for k,v in d['kernel_ms'].items()})" >> gpurun_out/r2_t15.log 2>&1
echo done
