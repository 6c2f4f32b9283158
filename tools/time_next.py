"""Time the kernels of the "next" rows on BASELINE-shaped batches, next to the CPU
oracle on a sample:   python tools/time_next.py
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from asr_dfcnn_transformer_b200 import _lib, ctc, features, noise, utils, wav_util  # noqa: E402
from oracle import fbank_ref, psf_ref, synth  # noqa: E402


def gpu_time(fn, n=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    lens = synth.ragged_lengths(rng, 256, 3.0, 7.0)
    audio_s = float(lens.sum()) / 16000

    # ---- 8f-2 mel front end, C2-shaped batch (device-resident float64 samples)
    sigs = [rng.standard_normal(int(n)) * 0.1 for n in lens]
    L = _lib.lib()
    counts = np.array([len(s) for s in sigs], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    nfr = np.array([wav_util.logfbank_frames(int(n)) for n in counts], dtype=np.int64)
    fo = np.concatenate([[0], np.cumsum(nfr)]).astype(np.int64)
    x_d = torch.from_numpy(np.concatenate(sigs)).to(dev)
    so, sc, fo_d = (torch.from_numpy(a).to(dev) for a in (offs[:-1].copy(), counts, fo))
    bins = torch.from_numpy(wav_util._mel_bins(200, 512, 16000)).to(dev)
    out = torch.empty((int(fo[-1]), 200), dtype=torch.float32, device=dev)

    def run_mel():
        st = L.asrk_logfbank_run(_lib.ptr(x_d), _lib.ptr(so), _lib.ptr(sc), _lib.ptr(fo_d), None, _lib.ptr(bins), 256,
                                 int(fo[-1]), 200, 400, 160, 0.97, 1, _lib.ptr(out), _lib.stream_ptr(None))
        assert st == 0
    ms = gpu_time(run_mel)

    def run_mel_raw():
        st = L.asrk_logfbank_run(_lib.ptr(x_d), _lib.ptr(so), _lib.ptr(sc), _lib.ptr(fo_d), None, _lib.ptr(bins), 256,
                                 int(fo[-1]), 200, 400, 160, 0.97, 0, _lib.ptr(out), _lib.stream_ptr(None))
        assert st == 0
    ms_raw = gpu_time(run_mel_raw)
    print("mel front end  : main kernel alone %.3f ms, z-score %.3f ms" % (ms_raw, ms - ms_raw))
    t0 = time.perf_counter()
    for s in sigs[:8]:
        psf_ref.compute_fbank_from_api(s)
    cpu = (time.perf_counter() - t0) / (counts[:8].sum() / 16000)
    print("mel front end  : %.3f ms per 256-utterance batch (%.0f audio-s) = %.2f M audio-s/s ; oracle %.1f audio-s/s/core"
          % (ms, audio_s, audio_s / ms / 1e3, 1 / cpu))

    # ---- 8f-1 coloured noise, C4-shaped batch of 512
    lens4 = synth.ragged_lengths(rng, 512, 3.0, 7.0)
    xs = [rng.standard_normal(int(n)) for n in lens4]
    cols = list(rng.integers(-10, 11, 512) / 10)
    ms = gpu_time(lambda: noise.color_noise_batch(xs, cols), n=2)      # includes the H2D of the deviates
    t0 = time.perf_counter()
    for x, c in zip(xs[:4], cols[:4]):
        fbank_ref.color_noise_from_normal(x, c)
    cpu = (time.perf_counter() - t0) / 4
    print("coloured noise : %.1f ms per 512-utterance batch incl. host->device (%.3f ms per utterance) ; oracle %.1f ms per utterance"
          % (ms, ms / 512, cpu * 1e3))

    # ---- 8f-3 / 8f-4
    feat = torch.randn((int(fo[-1]), 200), dtype=torch.float32, device=dev)
    ms = gpu_time(lambda: utils.lfr_batch(feat, fo, 4, 3))
    print("LFR stacking   : %.3f ms per batch (%.0f MB moved)" % (ms, feat.numel() * 4 * (1 + 4 / 3) / 1e6))
    hyp = torch.randint(0, 1400, (256, 90), dtype=torch.int32, device=dev)
    hl = torch.randint(5, 40, (256,), dtype=torch.int32, device=dev)
    tr = torch.randint(0, 1400, (256, 64), dtype=torch.int32, device=dev)
    tl = torch.randint(8, 25, (256,), dtype=torch.int32, device=dev)
    ms = gpu_time(lambda: utils.edit_distance(hyp, hl, tr, tl))
    print("label error    : %.3f ms per 256-utterance batch" % ms)


if __name__ == "__main__":
    main()
