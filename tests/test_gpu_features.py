"""GPU parity: spectrogram kernels vs the oracle restatement and vs the golden
vectors produced by the reference's own code (tools/make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from oracle import fbank_ref, synth
from tests.util import FEATURE_TOL, feature_err, zscore_feature_err

pytestmark = pytest.mark.gpu


def _gpu(signals, mode, **kw):
    from asr_dfcnn_transformer_b200 import features
    fb = features.compute_features(signals, mode=mode, **kw)
    return fb.features.cpu().numpy(), fb.frame_offsets


@pytest.mark.parametrize("mode,key", [("fbank", "fbank"), ("asrt", "asrt")])
def test_golden_vectors(golden_dir, mode, key):
    files = sorted(glob.glob(os.path.join(golden_dir, "fbank_*.npz")))
    assert files
    sigs, refs = [], []
    for f in files:
        d = np.load(f)
        sigs.append(d["pcm"])
        refs.append(d[key])
    out, fo = _gpu(sigs, mode)
    for i, f in enumerate(files):
        got = out[fo[i]:fo[i + 1]]
        assert got.shape == refs[i].shape, (f, got.shape, refs[i].shape)
        if mode == "fbank":
            err = zscore_feature_err(got, refs[i], fbank_ref.compute_fbank_unnormalised(sigs[i]))
        else:
            err = feature_err(got, refs[i])
        assert err <= FEATURE_TOL, (os.path.basename(f), mode, err)


def test_raw_vs_oracle_generators():
    rng = np.random.default_rng(7)
    sigs = [synth.g1_white(rng, 16000), synth.g2_voiced(rng, 40000), synth.g2_voiced(rng, 16080),
            synth.g1_white(rng, 401), synth.g1_white(rng, 300), synth.g2_voiced(rng, 5173)]
    sigs += list(synth.g3_edge_cases(rng).values())
    out, fo = _gpu(sigs, "fbank_raw")
    for i, s in enumerate(sigs):
        ref = fbank_ref.compute_fbank_unnormalised(s)
        got = out[fo[i]:fo[i + 1]]
        assert got.shape == ref.shape
        assert feature_err(got, ref) <= FEATURE_TOL, (i, feature_err(got, ref))


def test_short_last_frame_raises_like_reference():
    # N = 399: wav_util.py:61 still asks for one frame, and ``data_line * w`` then
    # fails to broadcast (399 vs 400) -> ValueError in the reference
    rng = np.random.default_rng(1)
    from asr_dfcnn_transformer_b200 import features
    with pytest.raises(ValueError):
        features.compute_features([synth.g1_white(rng, 399)], mode="fbank")
    with pytest.raises(ValueError):
        fbank_ref.compute_fbank(synth.g1_white(rng, 399))


def test_fbank_zscore_vs_oracle_ragged_batch():
    rng = np.random.default_rng(11)
    lens = synth.ragged_lengths(rng, 24, 0.5, 3.0)
    sigs = [synth.g2_voiced(rng, int(n)) if i % 2 else synth.g1_white(rng, int(n)) for i, n in enumerate(lens)]
    out, fo = _gpu(sigs, "fbank")
    worst = 0.0
    for i, s in enumerate(sigs):
        ref = fbank_ref.compute_fbank(s)
        worst = max(worst, feature_err(out[fo[i]:fo[i + 1]], ref))
    assert worst <= FEATURE_TOL, worst


def test_padded_loader_layout():
    rng = np.random.default_rng(5)
    sigs = [synth.g1_white(rng, 16000), synth.g1_white(rng, 9000)]
    from asr_dfcnn_transformer_b200 import features
    fb = features.compute_features(sigs, mode="fbank", padded_rows=1600)
    out = fb.features.cpu().numpy()
    assert out.shape == (2, 1600, 200)
    for i, s in enumerate(sigs):
        ref = fbank_ref.compute_fbank(s)
        assert feature_err(out[i, :ref.shape[0]], ref) <= FEATURE_TOL
        assert np.all(out[i, ref.shape[0]:] == 0)


def test_noise_golden(golden_dir):
    from asr_dfcnn_transformer_b200 import features, noise
    files = sorted(glob.glob(os.path.join(golden_dir, "noise_case*.npz")))
    assert files
    for f in files:
        d = np.load(f)
        K = noise.SNR2K(d["signal"], d["noise"], int(d["snr_db"]))
        assert K.dtype == np.float32
        assert K == d["K"], (f, K, d["K"])            # bit-exact gain
        mixed = noise.mix(d["signal"], d["noise"], int(d["snr_db"]))
        assert np.array_equal(mixed, d["mixed"])       # bit-exact mixed signal
        fb = features.compute_features([d["signal"]], noises=[d["noise"]], snr_db=[int(d["snr_db"])])
        err = feature_err(fb.features.cpu().numpy(), d["fbank"])
        assert err <= FEATURE_TOL, (f, err)


def test_wav_file_surface(tmp_path):
    import scipy.io.wavfile as wavfile
    from asr_dfcnn_transformer_b200 import wav_util
    rng = np.random.default_rng(3)
    sig = synth.g2_voiced(rng, 16240)
    p = str(tmp_path / "a.wav")
    wavfile.write(p, 16000, sig)
    got = wav_util.compute_fbank(p)
    assert got.dtype == np.float64
    assert feature_err(got, fbank_ref.compute_fbank(sig)) <= FEATURE_TOL
    got2 = wav_util.compute_fbank_from_asrt(p)
    assert feature_err(got2, fbank_ref.compute_fbank_from_asrt(sig)) <= FEATURE_TOL
    wd, fr = wav_util.read_wav_data(p)
    assert fr == 16000 and wd.shape == (1, 16240) and np.array_equal(wd[0], sig)


def test_loader_batch_assembly(tmp_path):
    """data_loader.py:105-162 on the device path: rejected rows are dropped, the kept
    rows land zero-padded in [B',1600,200,1] and match the oracle."""
    from asr_dfcnn_transformer_b200 import data_loader as dl
    d = tmp_path / "dict.txt"
    d.write_text("a1\tx\nb2\ty\nc3\tw\n", encoding="utf-8")
    _, s2i, _ = dl.load_acoustic_vocab(str(d))
    rng = np.random.default_rng(11)
    sigs = [synth.g2_voiced(rng, 24000), synth.g1_white(rng, 4000), synth.g1_white(rng, 16080),
            synth.g1_white(rng, 12345)]
    labs = ["a1 b2", "a1 b2 c3 a1", "c3", "b2 zz"]           # row 1: label >= T_ctc, row 3: unknown symbol
    wav, il, lab, ll, keep = dl.data_generation(sigs, labs, s2i, front_end="spectrogram")
    assert keep == [0, 2]
    assert tuple(wav.shape) == (2, 1600, 200, 1) and il.tolist() == [19, 13] and ll.tolist() == [2, 1]
    assert lab[0, :3].tolist() == [0, 1, 0] and lab[1, :2].tolist() == [2, 0]
    got = wav.cpu().numpy()[..., 0]
    for r, i in enumerate(keep):
        ref = fbank_ref.compute_fbank(sigs[i])
        n = ref.shape[0]
        err = zscore_feature_err(got[r, :n], ref, fbank_ref.compute_fbank_unnormalised(sigs[i]))
        assert err <= FEATURE_TOL, (i, err)
        assert not got[r, n:].any()


def test_batch_larger_than_one_launch_slice():
    """More utterances than the kernel's per-launch prefix array (2047): the C ABI slices
    the batch; every utterance must still come out right (1- and 2-frame utterances)."""
    rng = np.random.default_rng(3)
    base = [synth.g1_white(rng, 400), synth.g2_voiced(rng, 560), synth.g1_white(rng, 401)]
    sigs = [base[i % 3] for i in range(2100)]
    out, fo = _gpu(sigs, "fbank_raw")
    refs = [fbank_ref.compute_fbank_unnormalised(s) for s in base]
    assert fo[-1] == sum(r.shape[0] for r in refs) * 700
    for i in (0, 1, 2, 2045, 2046, 2047, 2048, 2049, 2099):
        assert feature_err(out[fo[i]:fo[i + 1]], refs[i % 3]) <= FEATURE_TOL, i
    out, fo = _gpu(sigs[:2050], "fbank")                     # z-scored: 1-frame utterances are all zeros
    for i in (0, 2046, 2047, 2049):
        ref = fbank_ref.compute_fbank(sigs[i])
        assert feature_err(out[fo[i]:fo[i + 1]], ref) <= FEATURE_TOL, i


def test_unaligned_sample_offsets_scalar_path():
    """Utterance starts that are not 16-byte aligned take the synchronous staging path."""
    import torch
    from asr_dfcnn_transformer_b200 import features
    rng = np.random.default_rng(4)
    a, b = synth.g2_voiced(rng, 8000), synth.g1_white(rng, 4321)
    buf = np.concatenate([np.zeros(3, np.int16), a, np.zeros(5, np.int16), b])
    so = np.array([3, 3 + len(a) + 5], dtype=np.int64)
    sc = np.array([len(a), len(b)], dtype=np.int64)
    nf = [features.n_frames_for(len(a)), features.n_frames_for(len(b))]
    fo = np.array([0, nf[0], nf[0] + nf[1]], dtype=np.int64)
    dev = torch.device("cuda")
    out = features.spectrogram_device(torch.from_numpy(buf).to(dev), torch.from_numpy(so).to(dev),
                                      torch.from_numpy(sc).to(dev), torch.from_numpy(fo).to(dev), 2, int(fo[-1]),
                                      "fbank").cpu().numpy()
    for i, s in enumerate((a, b)):
        ref = fbank_ref.compute_fbank(s)
        err = zscore_feature_err(out[fo[i]:fo[i + 1]], ref, fbank_ref.compute_fbank_unnormalised(s))
        assert err <= FEATURE_TOL, (i, err)


def test_full_size_c2_batch_properties():
    """BASELINE.json configs[1] at full size (256 utterances of 3..7 s): size-independent
    properties of the z-scored output -- every column of every utterance has mean 0 and
    standard deviation 1 (or is constant), and a spot-checked utterance matches the oracle."""
    import bench
    hb = bench.make_batch(2000)
    out, fo = _gpu(hb["pcm"], "fbank")
    assert fo[-1] == int(hb["nfr"].sum())
    for i in range(0, 256, 7):
        y = out[fo[i]:fo[i + 1]].astype(np.float64)
        assert np.abs(y.mean(axis=0)).max() < 2e-4
        sd = y.std(axis=0)
        assert np.all((np.abs(sd - 1.0) < 2e-3) | (sd == 0.0))
    i = 101
    ref = fbank_ref.compute_fbank(hb["pcm"][i])
    err = zscore_feature_err(out[fo[i]:fo[i + 1]], ref, fbank_ref.compute_fbank_unnormalised(hb["pcm"][i]))
    assert err <= FEATURE_TOL, err


def test_loader_batch_assembly_live_front_end(tmp_path):
    """The same rules with the front end the live loader calls (compute_fbank_from_api)."""
    from asr_dfcnn_transformer_b200 import data_loader as dl
    from oracle import psf_ref
    d = tmp_path / "dict.txt"
    d.write_text("a1\tx\nb2\ty\n", encoding="utf-8")
    _, s2i, _ = dl.load_acoustic_vocab(str(d))
    rng = np.random.default_rng(12)
    sigs = [synth.g2_voiced(rng, 24000).astype(np.float64) / 32768, synth.g1_white(rng, 8000).astype(np.float64) / 32768]
    wav, il, lab, ll, keep = dl.data_generation(sigs, ["a1 b2", "b2"], s2i)
    assert keep == [0, 1] and il.tolist() == [149 // 8 + 1, 49 // 8 + 1]
    got = wav.cpu().numpy()[..., 0]
    for r, s in enumerate(sigs):
        raw = psf_ref.logfbank(s)
        ref = psf_ref.compute_fbank_from_api(s)
        live = raw.std(axis=0) > 1e-9
        n = ref.shape[0]
        assert feature_err(got[r, :n][:, live], ref[:, live]) <= FEATURE_TOL
        assert not got[r, n:].any()


def test_c4_noise_augmented_batch():
    """BASELINE.json configs[3] shape (scaled to 96 utterances of ~2 s so that the float64 oracle
    stays in seconds): float32 signal + coloured noise at 5..10 dB, gains computed on the device,
    mix fused into the frame load, z-scored features vs the oracle of the mixed signal."""
    from asr_dfcnn_transformer_b200 import features
    rng = np.random.default_rng(4000)
    sigs, noises, dbs = [], [], []
    for i in range(96):
        n = int(rng.integers(24000, 40000))
        sigs.append((synth.g2_voiced(rng, n).astype(np.float32) / 32768.0))
        x = rng.standard_normal(n)                       # coloured noise stand-in (the generator is an input)
        nz = np.cumsum(x) if i % 2 else x
        nz = nz - nz.mean()
        noises.append((nz / nz.max()).astype(np.float32))
        dbs.append(int(rng.integers(5, 11)))
    fb = features.compute_features(sigs, noises=noises, snr_db=dbs)
    out = fb.features.cpu().numpy()
    worst = 0.0
    for i in range(0, 96, 5):
        s, nz = sigs[i], noises[i]
        K = np.float32(np.sqrt(np.mean(s ** 2) / np.mean(nz ** 2))) * np.float32(10 ** (-dbs[i] / 20))
        mixed = (s + np.float32(K) * nz).astype(np.float32)
        ref = fbank_ref.compute_fbank(mixed)
        got = out[fb.frame_offsets[i]:fb.frame_offsets[i + 1]]
        worst = max(worst, zscore_feature_err(got, ref, fbank_ref.compute_fbank_unnormalised(mixed)))
    assert worst <= 2 * FEATURE_TOL, worst      # the gain itself may differ by one float32 ulp from this numpy recipe
