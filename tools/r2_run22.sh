#!/bin/bash
# multi-GPU validation: the bench under torchrun at N = 4 and 8 (C2 headline), C5 at N = 8, reference arm at N = 8
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo.txt 2>&1
for n in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_b22_n$n.log 2>&1
  grep '^{' gpurun_out/r2_b22_n$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['results_stay_on_device']['value'], d['gpu_launches'])"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --workload c5 --steps 20 --warmup 5 > gpurun_out/r2_b22_c5_n8.log 2>&1
grep '^{' gpurun_out/r2_b22_c5_n8.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r2_b22_ref_n8.log 2>&1
grep '^{' gpurun_out/r2_b22_ref_n8.log | cut -c1-300
