import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from asr_dfcnn_transformer_b200 import _lib, ctc, features
what = sys.argv[1]
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
V = bench.V
db = bench.DeviceBatch(bench.make_batch(2000), dev, torch, "c2", "logits")
st = torch.cuda.Stream()
def body(stream):
    if what in ("main", "spec"):
        features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                    stream=stream, phases=_lib.PHASE_SPEC_SETUP | _lib.PHASE_SPEC_MAIN)
    if what in ("norm", "spec"):
        features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat,
                                    stream=stream, phases=_lib.PHASE_SPEC_NORMALIZE)
    if what == "specall":
        features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, "fbank", out=db.feat, stream=stream)
    if what == "ctc":
        return ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, V - 1, grad_scale=db.grad_scale,
                                 grad_out=db.grad, bounds=db.ctc_bounds, stream=stream)
with torch.cuda.stream(st):
    body(st)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=st):
    body(st)
torch.cuda.synchronize()
for it in range(3):
    g.replay()
    torch.cuda.synchronize()
print(what, "ok", float(db.feat.abs().sum()))
