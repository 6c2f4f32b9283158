"""Time the spectrogram MAIN kernel alone (CUDA events around the one launch) in three cache contexts:
  warm   three batches rotated back to back (what tools/time_spec.py measures: PCM partly L2 resident)
  flush  a 512 MB write between launches (L2 holds nothing of the batch)
  ctc    the fused CTC kernel of the same batch between launches (the step's real context)
    python tools/time_spec_ctx.py
"""
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
junk = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
V = bench.V


def spec(db, mode, phases):
    features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, mode, out=db.feat, phases=phases)


for mode, phases, label in (("fbank_raw", _lib.PHASE_ALL, "raw"), ("fbank", _lib.PHASE_SPEC_MAIN, "main+sums")):
    for ctx in ("warm", "flush", "ctc"):
        tot = 0.0
        n = 24
        for it in range(n + 3):
            db = pool[it % 3]
            if ctx == "flush":
                junk.fill_(it & 255)
            elif ctx == "ctc":
                ctc.ctc_loss_grad(db.logits, db.labels, db.label_len, db.input_len, V - 1, grad_scale=db.grad_scale,
                                  grad_out=db.grad, bounds=db.ctc_bounds)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            spec(db, mode, phases)
            e1.record()
            torch.cuda.synchronize()
            if it >= 3:
                tot += e0.elapsed_time(e1)
        print("%-10s %-6s %.1f us" % (label, ctx, 1e3 * tot / n))
