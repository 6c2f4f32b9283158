import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from oracle import build_c
from asr_dfcnn_transformer_b200 import ctc
hb = bench.make_batch(2001)
x, labels, ll, il = hb["logits"], hb["labels"], hb["label_len"], hb["input_len"]
V = x.shape[2]
r = ctc.ctc_loss_grad(torch.as_tensor(x).cuda(), labels, ll, il, V - 1, decode=(os.environ.get("DEC", "1") == "1"))
g = r.grad.cpu().numpy()
T = x.shape[0]
valid = np.arange(T)[:, None] < il[None, :]
s = np.abs(g.sum(-1)) * valid
bad = np.argwhere(s > 1e-4)
print("rows with |sum| > 1e-4:", len(bad), "max", s.max())
bs = sorted(set(bad[:, 1].tolist()))
print("utterances:", bs[:20])
for b in bs[:4]:
    ts = bad[bad[:, 1] == b][:, 0]
    print(" b", b, "T", il[b], "L", ll[b], "bad frames", ts.tolist()[:30])
    rl, rg, st = build_c.ctc_loss_grad(np.ascontiguousarray(x[:, b:b+1]), labels[b:b+1], ll[b:b+1], il[b:b+1], V - 1, real="f64")
    d = np.abs(g[:, b] - rg[:, 0])
    t = int(ts[0])
    v = int(np.argmax(d[t]))
    lab = labels[b, :ll[b]].tolist()
    print("   loss", float(r.loss[b]), rl[0], "max grad err at t", t, "class", v, d[t, v], "got", g[t, b, v], "ref", rg[t, 0, v], "is label pos", [j for j, c in enumerate(lab) if c == v], "blank" if v == V - 1 else "")
    print("   labels", lab)
