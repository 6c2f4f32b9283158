import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from asr_dfcnn_transformer_b200 import _lib, ctc, features
variant = sys.argv[1]
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
V = bench.V
db = bench.DeviceBatch(bench.make_batch(2000), dev, torch, "c2", "logits")
side = torch.cuda.Stream()
hi = torch.cuda.Stream(priority=-1 if "prio" in variant else 0)
ev_fork, ev_main, ev_j1, ev_j2 = (torch.cuda.Event() for _ in range(4))
def step():
    cur = torch.cuda.current_stream()
    ev_fork.record(cur)
    side.wait_event(ev_fork)
    if "gate" not in variant:
        hi.wait_event(ev_fork)
    with torch.cuda.stream(side):
        if "split" in variant:
            features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                        stream=side, phases=_lib.PHASE_SPEC_SETUP | _lib.PHASE_SPEC_MAIN)
            ev_main.record(side)
            features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                        stream=side, phases=_lib.PHASE_SPEC_NORMALIZE)
        else:
            features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat, stream=side)
        ev_j1.record(side)
    if "gate" in variant:
        hi.wait_event(ev_main)
    with torch.cuda.stream(hi):
        r = ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, V - 1, grad_scale=db.grad_scale,
                              grad_out=db.grad, bounds=db.ctc_bounds, stream=hi)
        ev_j2.record(hi)
    cur.wait_event(ev_j1)
    cur.wait_event(ev_j2)
    return r
step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    r = step()
torch.cuda.synchronize()
for it in range(3):
    g.replay()
    torch.cuda.synchronize()
print(variant, "ok", float(r.loss.sum()))
