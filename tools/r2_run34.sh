#!/bin/bash
# final check on two GPUs: the whole GPU suite (multi-rank test included), smoke, bench at N = 1 and N = 2, reference arm
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_t34.log 2>&1
tail -2 gpurun_out/r2_t34.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2_b34_n1.log 2>&1
grep '^{' gpurun_out/r2_b34_n1.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['steps'], d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['step_frac'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_b34_n2.log 2>&1
grep '^{' gpurun_out/r2_b34_n2.log | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print(d['n_gpus'], d['steps'], d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['results_stay_on_device']['value'], d['gpu_launches'])"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29652 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_b34_ref_n2.log 2>&1
grep '^{' gpurun_out/r2_b34_ref_n2.log | cut -c1-200
