"""GPU parity: coloured-noise generation (util/noise.py:17-34) vs the golden vectors made by the
reference's own color_noise and vs the oracle restatement, given the same normal deviates."""
import glob
import os

import numpy as np
import pytest

from oracle import fbank_ref

pytestmark = pytest.mark.gpu

TOL = 2e-6          # float32 output in [-a, 1]: a couple of ulps of 1


def test_golden_color_noise(golden_dir):
    from asr_dfcnn_transformer_b200 import noise
    files = sorted(glob.glob(os.path.join(golden_dir, "noise_case*.npz")))
    assert files
    for f in files:
        d = np.load(f)
        out, offs = noise.color_noise_batch([d["x_random"]], [float(d["colour"])])
        got = out.cpu().numpy()
        assert got.dtype == np.float32 and got.shape == d["noise"].shape
        assert np.abs(got - d["noise"]).max() <= TOL, (f, np.abs(got - d["noise"]).max())
        assert got.max() == 1.0                                   # divided by the maximum, not the abs-maximum


def test_ragged_batch_any_length_vs_oracle():
    from asr_dfcnn_transformer_b200 import noise
    rng = np.random.default_rng(5)
    lens = [2, 3, 5, 16, 17, 1000, 1001, 4096, 10007, 80000, 65536, 65537, 37123]
    cols = [0.0, -1.0, 1.0, 0.3, -0.7, 0.5, -0.5, 1.0, -1.0, 0.1, 0.9, -0.9, 0.0]
    xs = [rng.standard_normal(n) for n in lens]
    out, offs = noise.color_noise_batch(xs, cols)
    got = out.cpu().numpy()
    for i, (x, c) in enumerate(zip(xs, cols)):
        ref = fbank_ref.color_noise_from_normal(x, c)
        g = got[offs[i]:offs[i + 1]]
        assert np.abs(g - ref).max() <= TOL * max(1.0, np.abs(ref).max()), (lens[i], c, np.abs(g - ref).max())


@pytest.mark.parametrize("n", [700, 1500, 3000, 10000, 40001, 300000])
def test_every_transform_length_class(n):
    """One call per length so that the power-of-two transform length changes: 2^11 (one launch per
    radix-2 stage), 2^12 (one contiguous block), 2^13 / 2^15 / 2^17 (one strided pass of 1, 3, 5 stages:
    odd counts end with a single-stage round), 2^20 (two strided passes)."""
    from asr_dfcnn_transformer_b200 import noise
    rng = np.random.default_rng(n)
    x = rng.standard_normal(n)
    for c in (-1.0, 0.3):
        out, offs = noise.color_noise_batch([x], [c])
        ref = fbank_ref.color_noise_from_normal(x, c)
        g = out.cpu().numpy()
        assert np.abs(g - ref).max() <= TOL * max(1.0, np.abs(ref).max()), (n, c, np.abs(g - ref).max())


def test_drop_in_surface_reproduces_the_reference_stream():
    """Same np.random.seed -> the same draw as the reference's color_noise -> the same noise."""
    from asr_dfcnn_transformer_b200 import noise
    np.random.seed(1234)
    got = noise.color_noise(24001, -0.6)
    np.random.seed(1234)
    ref = fbank_ref.color_noise_from_normal(np.random.normal(0, 1, 24001), -0.6)
    assert got.dtype == np.float32 and np.abs(got - ref).max() <= TOL
    # and the full augmentation chain on the device: noise -> SNR2K -> mix
    sig = (0.3 * np.sin(np.arange(24001) * 0.05)).astype(np.float32)
    mixed = noise.mix(sig, got, 7)
    refm = fbank_ref.mix_noise(sig, got, 7)
    assert np.array_equal(mixed, refm)


def test_add_noise_in_memory_branch(tmp_path):
    """noise.py:70-128 with out_path=None: same random draws in the same order as the reference
    (random.randint for SNR and colour, np.random.normal for the deviates)."""
    import random
    import scipy.io.wavfile as wavfile
    from asr_dfcnn_transformer_b200 import noise
    t = np.arange(16000)
    pcm = np.round(9000 * np.sin(t * 0.03)).astype(np.int16)
    path = str(tmp_path / "a.wav")
    wavfile.write(path, 16000, pcm)
    random.seed(3)
    np.random.seed(3)
    out, names = noise.add_noise([path], n_to_add=2)
    assert names == [] and len(out) == 2 and out[0].dtype == np.float32 and out[0].shape == (16000,)
    random.seed(3)
    np.random.seed(3)
    sig = (pcm.astype(np.float32) / 32768.0).astype(np.float32)
    for k in range(2):
        snr = random.randint(5, 10)
        col = random.randint(-10, 10) / 10
        nz = fbank_ref.color_noise_from_normal(np.random.normal(0, 1, 16000), col)
        ref = fbank_ref.mix_noise(sig, nz, snr)
        assert np.abs(out[k] - ref).max() <= 1e-5
    assert noise.add_noise("/nonexistent/dir") == 0 and noise.add_noise([path], type_noise="3") == 0


def test_snr2k_bit_exact_every_tree_shape():
    """noise.py:48-52: the gain equals numpy's float32 pairwise sums bit for bit for every shape of the recursion tree
    the kernel distinguishes: plain loop (n < 8), one leaf, remainders, fewer leaves than one warp's sixteen, leaves
    above the bottom level (n just over a power of two), several passes per CTA, more than 2^20 samples; through the
    C ABI in ONE ragged batch, with 16-byte aligned and with unaligned utterance starts (scalar loads)."""
    import torch
    from asr_dfcnn_transformer_b200 import _lib
    rng = np.random.default_rng(48)
    lens = [1, 5, 8, 9, 127, 128, 129, 135, 255, 257, 1000, 1023, 1024, 1025, 2047, 2049, 4097, 16240, 65537, 79999,
            80000, 80001, 131072, 131073, 1048576 + 3, 3000001]
    sig = [(rng.standard_normal(n) * 0.1).astype(np.float32) for n in lens]
    noi = [rng.standard_normal(n).astype(np.float32) for n in lens]
    dbs = rng.integers(5, 11, len(lens)).astype(np.int32)
    for pad in (4, 3):                       # utterance starts: multiples of 4 floats / not
        offs = np.zeros(len(lens), dtype=np.int64)
        o = 0 if pad == 4 else 1
        for i, n in enumerate(lens):
            offs[i] = o
            o += (n + 3) // 4 * 4 + (0 if pad == 4 else 3)
        s = np.zeros(o + 8, np.float32)
        z = np.zeros(o + 8, np.float32)
        for i, n in enumerate(lens):
            s[offs[i]:offs[i] + n] = sig[i]
            z[offs[i]:offs[i] + n] = noi[i]
        dev = torch.device("cuda", 0)
        ds, dz = torch.as_tensor(s).to(dev), torch.as_tensor(z).to(dev)
        doff = torch.as_tensor(offs).to(dev)
        dcnt = torch.as_tensor(np.array(lens, dtype=np.int64)).to(dev)
        ddb = torch.as_tensor(dbs).to(dev)
        gain = torch.empty(len(lens), dtype=torch.float32, device=dev)
        rc = _lib.lib().asrk_snr2k_run(_lib.ptr(ds), _lib.ptr(dz), _lib.ptr(doff), _lib.ptr(dcnt), _lib.ptr(ddb),
                                       len(lens), _lib.ptr(gain), _lib.stream_ptr(None))
        _lib.check(rc, "asrk_snr2k_run")
        got = gain.cpu().numpy()
        for i, n in enumerate(lens):
            ref = fbank_ref.snr2k(sig[i], noi[i], int(dbs[i]))
            assert got[i] == np.float32(ref), (pad, n, got[i], ref)
