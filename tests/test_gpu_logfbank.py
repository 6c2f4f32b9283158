"""GPU parity: the mel filterbank front end (compute_fbank_from_api) vs the restated
python_speech_features algorithm (oracle/psf_ref.py; parity unpinned by the reference)."""
import numpy as np
import pytest

from oracle import psf_ref, synth
from tests.util import FEATURE_TOL, feature_err

pytestmark = pytest.mark.gpu


def _signals():
    rng = np.random.default_rng(21)
    sigs = [synth.g1_white(rng, 16000), synth.g2_voiced(rng, 40000), synth.g2_voiced(rng, 16080),
            synth.g1_white(rng, 401), synth.g1_white(rng, 300), synth.g2_voiced(rng, 5173)]
    return [s.astype(np.float64) / 32768.0 for s in sigs]          # what soundfile returns for 16-bit PCM


def test_raw_logfbank_vs_oracle():
    from asr_dfcnn_transformer_b200 import wav_util
    sigs = _signals()
    fb = wav_util.compute_fbank_from_api_batch(sigs, 16000, 200, normalise=False)
    out = fb.features.cpu().numpy()
    for i, s in enumerate(sigs):
        ref = psf_ref.logfbank(s)
        got = out[fb.frame_offsets[i]:fb.frame_offsets[i + 1]]
        assert got.shape == ref.shape, (i, got.shape, ref.shape)
        assert feature_err(got, ref) <= FEATURE_TOL, (i, feature_err(got, ref))


def test_zscored_vs_oracle_and_surface(tmp_path):
    from asr_dfcnn_transformer_b200 import wav_util
    sigs = _signals()[:3]
    for s in sigs:
        ref_raw = psf_ref.logfbank(s)
        ref = psf_ref.compute_fbank_from_api(s, 16000)
        got = wav_util.compute_fbank_from_api(s, 16000)
        assert got.dtype == np.float64 and got.shape == ref.shape
        live = ref_raw.std(axis=0) > 1e-9            # empty mel filters: the reference holds rounding noise there
        assert live.sum() == 157
        sd = ref_raw.std(axis=0)[live]
        cond = np.maximum(1.0, 0.05 / sd)             # same conditioning allowance as the spectrogram tests
        err = np.max(np.abs(got[:, live] - ref[:, live]) / np.maximum(np.abs(ref[:, live]), 1.0) / cond[None, :])
        assert err <= FEATURE_TOL, err
        assert not got[:, ~live].any()               # constant columns come out as exact zeros
    # the file surface (wav_util.py:13-19): int16 wav through read_wav_data; a constant gain of the
    # signal only shifts the log filterbank energies, which the z-score removes
    import scipy.io.wavfile as wavfile
    pcm = np.round(sigs[1] * 32768.0).astype(np.int16)
    path = str(tmp_path / "a.wav")
    wavfile.write(path, 16000, pcm)
    got = wav_util.compute_fbank_from_file(path, 200)
    # the oracle gets what the reference passes on: read_wav_data's [1, N] int16 array -- for which the
    # library's pre-emphasis expression leaves the samples unchanged (no pre-emphasis on this path)
    wd, fr = wav_util.read_wav_data(path)
    assert wd.shape == (1, len(pcm)) and fr == 16000
    ref = psf_ref.compute_fbank_from_api(wd, 16000)
    ref_raw = psf_ref.logfbank(wd)
    assert np.array_equal(psf_ref.preemphasis(np.asarray(wd, dtype=np.float64)), pcm.astype(np.float64))
    live = ref_raw.std(axis=0) > 1e-9
    assert feature_err(got[:, live], ref[:, live]) <= FEATURE_TOL
    # soundfile route (sf_flag=True): 1-D float64 in [-1, 1], pre-emphasis 0.97 applies
    got = wav_util.compute_fbank_from_file(path, 200, sf_flag=True)
    ref = psf_ref.compute_fbank_from_api(pcm.astype(np.float64) / 32768.0, 16000)
    assert feature_err(got[:, live], ref[:, live]) <= FEATURE_TOL


@pytest.mark.parametrize("fs,nfilt", [(8000, 40), (16000, 26), (20000, 200), (11025, 64)])
def test_other_sample_rates_and_filter_counts(fs, nfilt):
    """frame length 200 / 400 / 500 / 276 samples (round-half-up of 25 ms) inside the 512-point
    transform, and filter banks that are not 200 wide: the kernel's group counts, bin tables and lane
    loops are all derived from these."""
    from asr_dfcnn_transformer_b200 import wav_util
    rng = np.random.default_rng(fs + nfilt)
    sigs = [synth.g2_voiced(rng, 3 * fs).astype(np.float64) / 32768.0,
            synth.g1_white(rng, fs // 2 + 7).astype(np.float64) / 32768.0]
    fb = wav_util.compute_fbank_from_api_batch(sigs, fs, nfilt, normalise=False)
    out = fb.features.cpu().numpy()
    for i, s in enumerate(sigs):
        ref = psf_ref.logfbank(s, fs, nfilt)
        got = out[fb.frame_offsets[i]:fb.frame_offsets[i + 1]]
        assert got.shape == ref.shape, (i, got.shape, ref.shape)
        assert feature_err(got, ref) <= FEATURE_TOL, (fs, nfilt, i, feature_err(got, ref))
