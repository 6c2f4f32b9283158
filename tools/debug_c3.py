import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import synth, build_c
from asr_dfcnn_transformer_b200 import ctc
rng = np.random.default_rng(3001)
V = synth.VOCAB_DICT_TXT
il = np.full(64, 1998, dtype=np.int32)
il[5], il[17] = 1500, 1001
x, labels, ll, il = synth.ctc_batch(rng, il, V, 280, 320)
xs = torch.as_tensor(x).cuda()
for rep in range(3):
    r = ctc.ctc_loss_grad(xs, labels, ll, il, V - 1, decode=True)
    s = torch.abs(r.grad.sum(-1)).cpu().numpy()
    T = x.shape[0]
    valid = np.arange(T)[:, None] < il[None, :]
    s = s * valid
    bad = np.argwhere(s > 1e-4)
    print("rep", rep, "bad rows", len(bad), "max", s.max(), "utterances", sorted(set(bad[:, 1].tolist()))[:20])
    for b in sorted(set(bad[:, 1].tolist()))[:3]:
        ts = bad[bad[:, 1] == b][:, 0]
        print("  b", b, "T", il[b], "L", ll[b], "n bad", len(ts), "frames", ts.tolist()[:16], "...", ts.tolist()[-4:])
        t = int(ts[0])
        g = r.grad[t, b].cpu().numpy()
        v = int(np.argmax(np.abs(g)))
        lab = labels[b, :ll[b]].tolist()
        print("    largest |g| at class", v, g[v], "label positions", [j for j, c in enumerate(lab) if c == v][:6], "blank" if v == V - 1 else "")
b = 0
rl, rg, st = build_c.ctc_loss_grad(np.ascontiguousarray(x[:, b:b+1]), labels[b:b+1], ll[b:b+1], il[b:b+1], V - 1, real="f64")
g = r.grad[:, b].cpu().numpy()
print("loss", float(r.loss[b]), rl[0])
lab = labels[b, :ll[b]].tolist()
for t in (1057, 1058, 1061, 1065, 1066):
    d = np.abs(g[t] - rg[t, 0])
    idx = np.argsort(-d)[:4]
    print(" t", t, "rowsum", g[t].sum(), [(int(v), float(g[t, v]), float(rg[t, 0, v]), [j for j, c in enumerate(lab) if c == v][:4]) for v in idx])
