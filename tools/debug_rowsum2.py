import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from asr_dfcnn_transformer_b200 import ctc
for seed in (2001, 2000, 2002):
    hb = bench.make_batch(seed)
    x, labels, ll, il = hb["logits"], hb["labels"], hb["label_len"], hb["input_len"]
    V = x.shape[2]
    xs = torch.as_tensor(x).cuda()
    T = x.shape[0]
    valid = torch.as_tensor(np.arange(T)[:, None] < il[None, :]).cuda()
    worst = 0.0
    nbad = 0
    ref = None
    for it in range(20):
        r = ctc.ctc_loss_grad(xs, labels, ll, il, V - 1, decode=True)
        s = (r.grad.sum(-1).abs() * valid)
        worst = max(worst, float(s.max()))
        nbad += int((s > 1e-4).sum())
        if ref is None:
            ref = r.grad.clone()
        elif not torch.equal(ref, r.grad):
            nbad += 1000000
    print("seed", seed, "20 runs: worst |row sum|", worst, "bad rows", nbad)
