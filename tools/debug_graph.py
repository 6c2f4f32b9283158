import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from asr_dfcnn_transformer_b200 import pipeline
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
V = bench.V
pool = [bench.DeviceBatch(bench.make_batch(2000 + i), dev, torch, "c2", "logits") for i in range(2)]
hp = pipeline.HotPathStep(dev)
acc = torch.zeros(2, dtype=torch.float64, device=dev)
gs = []
for db in pool:
    g, f, r = hp.capture(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, db.logits, db.labels, db.label_len,
                         db.input_len, V - 1, feat_out=db.feat, grad_out=db.grad, grad_scale=db.grad_scale,
                         ctc_bounds=db.ctc_bounds, loss_acc=acc if os.environ.get("ACC", "1") == "1" else None)
    gs.append((g, r))
    torch.cuda.synchronize()
    print("captured")
acc.zero_()
for it in range(6):
    gs[it % 2][0].replay()
    torch.cuda.synchronize()
    print("replay", it, float(gs[it % 2][1].loss.sum()), acc.tolist())
