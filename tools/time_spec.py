"""Time the spectrogram kernel alone on one C2 batch: z-scored vs raw output.
    python tools/time_spec.py [batch]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
from asr_dfcnn_transformer_b200 import features  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
pool = [bench.DeviceBatch(bench.make_batch(2000 + i, batch=batch), dev, torch, "c2", "logits") for i in range(3)]
for mode in os.environ.get("ASRK_TIME_MODES", "fbank,fbank_raw,asrt").split(","):
    for it in range(3):
        db = pool[it % 3]
        features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, mode, out=db.feat)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for it in range(n):
        db = pool[it % 3]
        features.spectrogram_device(db.samples, db.so, db.sc, db.fo, db.B, db.total_frames, mode, out=db.feat)
    e1.record()
    torch.cuda.synchronize()
    print("%-10s %.1f us per batch (%d frames)" % (mode, 1e3 * e0.elapsed_time(e1) / n, pool[0].total_frames))
