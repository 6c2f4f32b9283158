"""CPU: structural known answers of the restated python_speech_features front end."""
import numpy as np

from oracle import psf_ref


def test_frame_count_and_padding():
    assert psf_ref.num_frames(160000) == 999            # 10 s: one more than the in-repo spectrogram's 998
    assert psf_ref.num_frames(400) == 1 and psf_ref.num_frames(401) == 2 and psf_ref.num_frames(100) == 1
    fr = psf_ref.framesig(np.arange(1, 562, dtype=np.float64), 400, 160)
    assert fr.shape == (3, 400) and fr[2, 240] == 561 and fr[2, 241] == 0      # zero padded tail


def test_filterbank_structure():
    fb = psf_ref.get_filterbanks(200, 512, 16000)
    assert fb.shape == (200, 257)
    assert int((fb.sum(axis=1) == 0).sum()) == 43        # 43 of the 200 filters are empty
    assert int((fb != 0).sum()) == 353 and fb.max() <= 1.0
    b = psf_ref.mel_bins(200, 512, 16000)
    assert b[0] == 0 and b[-1] == 256 and np.all(np.diff(b) >= 0)


def test_logfbank_on_a_tone():
    t = np.arange(16000) / 16000.0
    x = 0.5 * np.sin(2 * np.pi * 1000 * t)
    f = psf_ref.logfbank(x)
    assert f.shape == (99, 200)
    eps_col = np.log(np.finfo(float).eps)
    assert np.isclose(f[:, fb_empty()].max(), eps_col) and np.isclose(f[:, fb_empty()].min(), eps_col)
    # the 43 empty filters give constant columns log(eps): after sklearn's scale they hold nothing but
    # float64 rounding noise of the mean (0 or +-O(1) depending on the frame count) -- parity is only
    # defined on the other columns (tests/test_gpu_logfbank.py masks columns with std < 1e-9)
    assert f[:, fb_empty()].std(axis=0).max() < 1e-12
    z = psf_ref.compute_fbank_from_api(x)
    live = np.setdiff1d(np.arange(200), fb_empty())
    assert np.abs(z[:, live].mean(axis=0)).max() < 1e-9 and np.abs(z[:, live].std(axis=0) - 1).max() < 1e-9


def fb_empty():
    return np.where(psf_ref.get_filterbanks(200, 512, 16000).sum(axis=1) == 0)[0]
