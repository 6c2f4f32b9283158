"""The step's tail in isolation: z-score kernels (stats + normalize) and the fused CTC kernel of a C2 batch, alone and
together on two streams (equal priorities in both launch orders, CTC on a high-priority stream)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from asr_dfcnn_transformer_b200 import _lib, ctc, features  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
pool = [bench.DeviceBatch(bench.make_batch(2000 + i), dev, torch, "c2", "logits") for i in range(3)]
V = bench.V
s_norm = torch.cuda.Stream()
s_ctc = torch.cuda.Stream()
s_hi = torch.cuda.Stream(priority=-1)
for db in pool:   # raw rows + unit sums in place
    features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                phases=_lib.PHASE_SPEC_SETUP | _lib.PHASE_SPEC_MAIN)
torch.cuda.synchronize()


def norm(db, st):
    features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                phases=_lib.PHASE_SPEC_NORMALIZE, stream=st)


def fused(db, st):
    ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, V - 1, grad_scale=db.grad_scale, grad_out=db.grad,
                      bounds=db.ctc_bounds, stream=st)


def one(mode, db, cur):
    fork, j1, j2 = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
    fork.record(cur)
    if mode == "norm":
        s_norm.wait_event(fork); norm(db, s_norm); j1.record(s_norm); cur.wait_event(j1)
    elif mode == "ctc":
        s_ctc.wait_event(fork); fused(db, s_ctc); j1.record(s_ctc); cur.wait_event(j1)
    else:
        sc = s_hi if "hi" in mode else s_ctc
        s_norm.wait_event(fork); sc.wait_event(fork)
        if "ctcfirst" in mode:
            fused(db, sc); norm(db, s_norm)
        else:
            norm(db, s_norm); fused(db, sc)
        j1.record(s_norm); j2.record(sc); cur.wait_event(j1); cur.wait_event(j2)


def run(mode, n=30):
    # one CUDA graph per batch of the pool (no host enqueue time inside the timed region)
    for db in pool:
        for st in (s_norm, s_ctc, s_hi):
            norm(db, st); fused(db, st)          # workspaces of every stream exist before the capture
    torch.cuda.synchronize()
    graphs = []
    for db in pool:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            one(mode, db, torch.cuda.current_stream())
        graphs.append(g)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        graphs[it % 3].replay()
    e0.record()
    for it in range(n):
        graphs[it % 3].replay()
    e1.record()
    torch.cuda.synchronize()
    print("%-22s %.1f us (graph replay)" % (mode, 1e3 * e0.elapsed_time(e1) / n))


for m in ("norm", "ctc", "both_normfirst", "both_ctcfirst", "both_hi_normfirst", "both_hi_ctcfirst"):
    run(m)
