#!/bin/bash
# two GPUs: default bench (two steps in flight, all-reduce on the side stream)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29744 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b44_n2.log 2>&1
grep '^{' gpurun_out/r2_b44_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['steps'], d['ms_per_step'], d['value'], d['loss_mean'], d['e2e']['value'], d['e2e']['results_stay_on_device']['value'], d['gpu_launches'], d['config']['steps_in_flight'])"
tail -3 gpurun_out/r2_b44_n2.log | cut -c1-200
