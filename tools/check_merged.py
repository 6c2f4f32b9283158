"""The merged tail (fused CTC kernel + z-score co-work) on a C2 batch: features after an eager step and after a CUDA
graph replay equal the two-kernel result bit for bit; step time eager / replayed / two-kernel; the merged kernel alone."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from asr_dfcnn_transformer_b200 import _lib, ctc, features, pipeline  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
pool = [bench.DeviceBatch(bench.make_batch(2000 + i, 256, "c2"), dev, torch, "c2", "logits") for i in range(3)]
V = bench.V
refs = []
for db in pool:
    refs.append(features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank").clone())
torch.cuda.synchronize()


def timeit(fn, n=30):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


for merged in (True, False):
    hp = pipeline.HotPathStep(dev, merged_tail=merged)
    for db in pool:
        hp.reserve(db.B, db.total_frames, db.logits.shape[0], db.labels.shape[1])

    def step(i, hp=hp):
        db = pool[i % 3]
        return hp(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, db.logits, db.labels, db.label_len, db.input_len,
                  V - 1, feat_out=db.feat, grad_out=db.grad, grad_scale=db.grad_scale, ctc_bounds=db.ctc_bounds)
    for i in range(3):
        f, r = step(i)
        torch.cuda.synchronize()
        assert torch.equal(f, refs[i]), ("eager", merged, i)
    t_eager = timeit(step)
    graphs = []
    for i in range(3):
        db = pool[i]
        g, f, r = hp.capture(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, db.logits, db.labels, db.label_len,
                             db.input_len, V - 1, feat_out=db.feat, grad_out=db.grad, grad_scale=db.grad_scale,
                             ctc_bounds=db.ctc_bounds)
        graphs.append(g)
    for i in range(3):
        pool[i].feat.zero_()
        graphs[i].replay()
        torch.cuda.synchronize()
        assert torch.equal(pool[i].feat[:pool[i].total_frames], refs[i]), ("replay", merged, i)
    t_graph = timeit(lambda i: graphs[i % 3].replay())
    print("merged tail %-5s  eager %.1f us  graph replay %.1f us  (features bit-identical to the two-kernel call)" % (
        merged, t_eager, t_graph), flush=True)

# the merged kernel alone, behind transform + statistics (same stream), and the pieces
def tail_only(i, merged):
    db = pool[i % 3]
    if merged:
        zw = features.zscore_work(db.feat, db.fo, db.B, db.total_frames)
        ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, V - 1, grad_scale=db.grad_scale, grad_out=db.grad,
                          bounds=db.ctc_bounds, zscore=zw)
    else:
        features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                    phases=_lib.PHASE_SPEC_NORMALIZE)
        ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, V - 1, grad_scale=db.grad_scale, grad_out=db.grad,
                          bounds=db.ctc_bounds)


def front(i):
    db = pool[i % 3]
    features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                phases=_lib.PHASE_SPEC_SETUP | _lib.PHASE_SPEC_MAIN | _lib.PHASE_SPEC_STATS)


t_front = timeit(front)
for merged in (True, False):
    t = timeit(lambda i: (front(i), tail_only(i, merged)))
    print("one stream: transform + statistics %.1f us, + tail (merged=%s) %.1f us -> tail %.1f us" % (t_front, merged, t, t - t_front))
