"""The z-score co-work alone: a one-utterance CTC batch (its CTA finishes in microseconds) carrying the z-score of a
full C2 feature batch, against the standalone z-score kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import bench  # noqa: E402
from asr_dfcnn_transformer_b200 import _lib, ctc, features  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
pool = [bench.DeviceBatch(bench.make_batch(2000 + i, 256, "c2"), dev, torch, "c2", "logits") for i in range(3)]
V = bench.V
x1 = torch.randn(4, 1, V, device=dev)
lab1 = torch.zeros((1, 8), dtype=torch.int32, device=dev)
ll1 = torch.ones(1, dtype=torch.int32, device=dev)
il1 = torch.full((1,), 4, dtype=torch.int32, device=dev)


def front(db):
    features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                phases=_lib.PHASE_SPEC_SETUP | _lib.PHASE_SPEC_MAIN | _lib.PHASE_SPEC_STATS)


def timed(name, fn, n=12):
    ts = []
    for i in range(n + 2):
        db = pool[i % 3]
        front(db)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(400000)
        e0.record()
        fn(db)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(1e3 * e0.elapsed_time(e1))
    print("%-50s %.1f us (min %.1f)" % (name, float(np.mean(ts)), float(np.min(ts))), flush=True)


def zonly(db):
    zw = features.zscore_work(db.feat, db.fo, db.B, db.total_frames)
    ctc.ctc_loss_grad(x1, lab1, ll1, il1, V - 1, bounds=(4, 1), zscore=zw)


def znorm(db):
    features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                phases=_lib.PHASE_SPEC_NORMALIZE)


def merged(db):
    zw = features.zscore_work(db.feat, db.fo, db.B, db.total_frames)
    ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, V - 1, grad_scale=db.grad_scale, grad_out=db.grad,
                      bounds=db.ctc_bounds, zscore=zw)


def ctc_only(db):
    ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, V - 1, grad_scale=db.grad_scale, grad_out=db.grad,
                      bounds=db.ctc_bounds)


timed("z-score co-work alone (296 CTAs)", zonly)
timed("standalone stats + z-score kernels", znorm)
timed("fused CTC alone", ctc_only)
timed("merged: fused CTC + z-score co-work", merged)
