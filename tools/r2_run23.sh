#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for n in 2 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n tools/pcie_ceiling_multi.py > gpurun_out/r2_pcie_multi_n$n.log 2>&1
grep '^N=' gpurun_out/r2_pcie_multi_n$n.log
done
ASRK_BENCH_LOGITS_IN=zero_copy timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_b23_n8_zc.log 2>&1
grep '^{' gpurun_out/r2_b23_n8_zc.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['results_stay_on_device']['value'])"
