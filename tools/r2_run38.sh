#!/bin/bash
# final multi-GPU check of the shipped bench: N = 8 and N = 4, C2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2970$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_b38_n$n.log 2>&1
  grep '^{' gpurun_out/r2_b38_n$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['results_stay_on_device']['value'], d['e2e']['what'][60:110], d['gpu_launches'])"
done
