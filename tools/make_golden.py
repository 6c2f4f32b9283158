"""Generate tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):
    python tools/make_golden.py

Each vector holds the input signal and the float64 output of the unmodified
reference function imported through oracle/ref_import.py:
  util.wav_util.compute_fbank            (wav_util.py:49-79)
  util.wav_util.compute_fbank_from_asrt  (wav_util.py:82-112)
  util.noise.SNR2K / color_noise / the mix of noise.py:108
The reference functions take wav *paths*; the signals are written to a temp dir
with scipy.io.wavfile.write (int16, or float32 for the noise-mixed case).
CTC / greedy-decode vectors are NOT produced here: TensorFlow/Keras cannot be
installed, so those oracles are unpinned against the reference.
"""
import os
import random
import sys
import tempfile

import numpy as np
import scipy.io.wavfile as wavfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, synth  # noqa: E402


def ref_fbank(wav_util, sig, which="fbank"):
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "x.wav")
        wavfile.write(p, 16000, sig)
        if which == "fbank":
            return wav_util.compute_fbank(p)
        return wav_util.compute_fbank_from_asrt(p)


def main():
    wav_util, noise = ref_import.load()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(20261018)

    cases = {
        "g1_white_1s": synth.g1_white(rng, 16000),
        "g2_voiced_1p5s": synth.g2_voiced(rng, 24000),
        "g2_voiced_hazard_16080": synth.g2_voiced(rng, 16080),
        "g2_voiced_hazard_16240": synth.g2_voiced(rng, 16240),
        "click": synth.g3_edge_cases(rng)["click"][:8000],
        "square_p74": synth.g3_edge_cases(rng)["square_p74"][:8000],
        "silence": np.zeros(4000, dtype=np.int16),
        "min_length_400": synth.g1_white(rng, 400),
    }
    for name, sig in cases.items():
        fb = ref_fbank(wav_util, sig, "fbank")
        asrt = ref_fbank(wav_util, sig, "asrt")
        np.savez_compressed(os.path.join(out_dir, f"fbank_{name}.npz"),
                            pcm=sig, fbank=fb, asrt=asrt)
        print(name, sig.shape, fb.shape, asrt.shape)

    # noise: reference color_noise (global numpy RNG, noise.py:18), SNR2K, mix,
    # then compute_fbank of the float32 mixed signal.
    for i, (colour, db) in enumerate([(0.0, 5), (-1.0, 10), (0.7, 8)]):
        n = 12000 + 37 * i  # even and odd lengths (noise.py:24-27)
        sig = (synth.g2_voiced(rng, n).astype(np.float32) / np.float32(32768.0)).astype(np.float32)
        np.random.seed(1234 + i)
        x_random_state = np.random.get_state()
        nz = noise.color_noise(n, colour)
        np.random.set_state(x_random_state)
        x_random = np.random.normal(0, 1, n)      # the same draw color_noise consumed
        K = noise.SNR2K(sig, nz, db)
        mixed = (sig + K * nz).astype(np.float32)  # noise.py:108
        fb = ref_fbank(wav_util, mixed, "fbank")
        np.savez_compressed(os.path.join(out_dir, f"noise_case{i}.npz"),
                            signal=sig, x_random=x_random, noise=nz, colour=colour, snr_db=db,
                            K=np.asarray(K), K_dtype=str(np.asarray(K).dtype),
                            mixed=mixed, fbank=fb)
        print("noise", i, n, colour, db, float(K), np.asarray(K).dtype, fb.shape)


if __name__ == "__main__":
    main()
