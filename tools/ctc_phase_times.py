"""Debug build only (ASRK_EXTRA_NVCC=-DASRK_CTC_TIMING): per-CTA phase times of the fused CTC kernel
on one C2 batch, read back from the token rows.
    ASRK_EXTRA_NVCC=-DASRK_CTC_TIMING python -m asr_dfcnn_transformer_b200._build && python tools/ctc_phase_times.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from asr_dfcnn_transformer_b200 import ctc  # noqa: E402

dev = torch.device("cuda", 0)
db = bench.DeviceBatch(bench.make_batch(2000), dev, torch, "c2", "logits")
for it in range(4):
    r = ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, bench.V - 1, grad_scale=db.grad_scale,
                          grad_out=db.grad, decode=True, bounds=db.ctc_bounds)
torch.cuda.synchronize()
t = r.tokens[:, :14].cpu().numpy().astype(np.int64)
GHZ = 1.965
for c_ in (1, 2, 3, 4, 8, 9, 10, 11, 12, 13):
    t[:, c_] = (t[:, c_] / GHZ).astype(np.int64)      # SM cycles -> ns
t[:, 0] = 0
t0 = t[:, 0] - t[:, 0].min()
t0[t0 < 0] += 1 << 30
A, B, P, C = t[:, 1], t[:, 2] - t[:, 1], t[:, 3] - t[:, 2], t[:, 4] - t[:, 3]
end = t0 + t[:, 4]
print("kernel span (first start -> last end) %.1f us; starts spread %.1f us" % (end.max() / 1e3, t0.max() / 1e3))
for name, v in (("A rows", A), ("B lattice", B), ("prefix", P), ("C grad", C), ("total", t[:, 4])):
    print("%-10s mean %6.1f  min %6.1f  max %6.1f us" % (name, v.mean() / 1e3, v.min() / 1e3, v.max() / 1e3))
print("prefix split: scan %.1f  logp %.1f  sK+barrier %.1f us" % (t[:, 8].mean() / 1e3, t[:, 9].mean() / 1e3, (P - t[:, 8] - t[:, 9]).mean() / 1e3))
per_T = np.polyfit(t[:, 6], B, 1)
print("B per frame: %.0f ns (+%.0f)" % (per_T[0], per_T[1]))
print("A per row-round (T/8): %.0f ns" % (np.polyfit(np.ceil(t[:, 6] / 8), A, 1)[0]))
print("C per row-round: %.0f ns" % (np.polyfit(np.ceil(t[:, 6] / 8), C, 1)[0]))
sm = t[:, 5]
cnt = np.bincount(sm, minlength=148)
busy = np.zeros(148)
for s_ in range(148):
    m = sm == s_
    if m.any():
        busy[s_] = end[m].max() / 1e3
print("CTAs per SM: 1 on %d SMs, 2 on %d SMs; SM finish time mean %.1f max %.1f us; single-CTA SMs finish %.1f, pairs %.1f"
      % ((cnt == 1).sum(), (cnt == 2).sum(), busy[cnt > 0].mean(), busy.max(), busy[cnt == 1].mean() if (cnt == 1).any() else 0,
         busy[cnt == 2].mean() if (cnt == 2).any() else 0))
idx = np.argsort(-end)[:5]
print("last CTAs: ", [(int(b), int(t[b, 6]), int(t[b, 7]), round(t0[b] / 1e3, 1), round(end[b] / 1e3, 1)) for b in idx])
pairs = {}
for b_ in range(len(sm)):
    pairs.setdefault(int(sm[b_]), []).append(b_)
two = [v for v in pairs.values() if len(v) == 2]
print("second CTA index minus first on shared SMs:", sorted(set(v[1] - v[0] for v in two)))
print("block indices alone on an SM:", sorted(v[0] for v in pairs.values() if len(v) == 1)[:50])
print("tick5 - tick2 values:", np.unique(t[:, 8])[:20], "quantiles", np.percentile(t[:, 8], [5, 50, 95]))
print("all deltas mod 32:", np.unique(np.concatenate([t[:, 1], t[:, 2], t[:, 3], t[:, 4]]) % 32)[:10])
print("tick7 - tick5 quantiles (ns)", np.percentile(t[:, 10], [5, 50, 95]))
for nm, c_ in (("alpha", 11), ("beta", 12), ("collapse", 13)):
    fit = np.polyfit(t[:, 6], t[:, c_], 1)
    print("%-9s mean %.1f us, per frame %.0f ns (+%.0f)" % (nm, t[:, c_].mean() / 1e3, fit[0], fit[1]))
