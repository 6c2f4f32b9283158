import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import ctc_ref, synth
from asr_dfcnn_transformer_b200 import ctc
rng = np.random.default_rng(3)
for T, V, L in ((50, 40, 40), (50, 40, 70), (200, 40, 100)):
    x, labels, ll, il = synth.ctc_batch(rng, [T], V, L, L)
    r = ctc.ctc_loss_grad(torch.as_tensor(x).cuda(), labels, ll, il, V - 1)
    rl, rg, ok = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, V - 1)
    g = r.grad.cpu().numpy()
    d = np.abs(g - rg)[:, 0]
    bad = np.argwhere(d > 1e-3)
    print("T", T, "L", int(ll[0]), "loss", float(r.loss[0]), rl[0], "bad", len(bad), "status", r.row_status.cpu().numpy())
    lab = labels[0, :ll[0]]
    ts = sorted(set(bad[:, 0].tolist()))
    print(" bad frames", ts[:40])
    for t, v in bad[:12]:
        pos = [j for j in range(len(lab)) if lab[j] == v]
        print("  t", t, "class", v, "label positions", pos, "got", g[t, 0, v], "ref", rg[t, 0, v])
