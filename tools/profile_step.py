"""Small driver for ncu: builds one C2 batch and runs a few steps of the hot path.
    python tools/profile_step.py [steps] [workload] [surface] [merged]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
workload = sys.argv[2] if len(sys.argv) > 2 else "c2"
surface = sys.argv[3] if len(sys.argv) > 3 else "logits"
bench.MERGED_TAIL = len(sys.argv) > 4 and sys.argv[4] == "merged"      # the z-score as co-work of the fused CTC kernel
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
w = bench.WORKLOADS[workload]
db = bench.DeviceBatch(bench.make_batch(2000, w["batch"], workload), dev, torch, workload, surface)
for _ in range(steps):
    r = bench.run_step(db, surface)
torch.cuda.synchronize()
if r is not None:
    print("loss sum", float(r.loss.sum()), "status max", int(r.row_status.max()))
