"""CPU restatement (numpy, float64) of the reference spectrogram features + noise mix.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never imported by the
product package.

Follows /root/reference/util/wav_util.py:49-79 (``compute_fbank``),
:82-112 (``compute_fbank_from_asrt``) and /root/reference/util/noise.py:48-52,
:108 (``SNR2K`` and the mix) statement by statement, but takes the already
decoded sample array instead of a wav path (wav decoding is file I/O, outside
the hot path).  Pinned against the reference's own code by
tests/test_oracle_features.py (live import in the build container) and by the
golden vectors under tests/golden/.
"""
import numpy as np
from scipy.fftpack import fft

FRAME_LEN = 400      # wav_util.py:51  (25 ms at 16 kHz)
FRAME_HOP = 160      # wav_util.py:67  (10 ms at 16 kHz)
N_BINS = 200         # wav_util.py:75


def hamming400():
    """wav_util.py:51-52 -- symmetric Hamming over 400 points, float64."""
    x = np.linspace(0, 400 - 1, 400, dtype=np.int64)
    return 0.54 - 0.46 * np.cos(2 * np.pi * (x) / (400 - 1))


def n_frames_fbank(n_samples: int, fs: int = 16000) -> int:
    """wav_util.py:61 -- evaluated as the *same Python float expression*.

    (the pre-emphasised signal of :59 has the same length as the input)"""
    return int(n_samples / fs * 1000 - 25) // 10 + 1


def n_frames_asrt(n_samples: int, fs: int = 16000) -> int:
    """wav_util.py:96 -- no '+1'."""
    return int(n_samples / fs * 1000 - 25) // 10


def zscore_columns(feature):
    """sklearn.preprocessing.scale(X) as called at wav_util.py:79 (axis=0,
    with_mean, with_std), restated: nanmean / nanstd(ddof=0), scales below
    10*eps replaced by 1, re-centring when the centred mean is not ~0."""
    X = np.array(feature, dtype=np.float64)
    if X.shape[0] == 0:
        return X
    mean_ = np.nanmean(X, 0)
    scale_ = np.nanstd(X, 0)
    X -= mean_
    mean_1 = np.nanmean(X, axis=0)
    if not np.allclose(mean_1, 0):
        X -= mean_1
    scale_ = scale_.copy()
    scale_[scale_ < 10 * np.finfo(scale_.dtype).eps] = 1.0
    X /= scale_
    mean_2 = np.nanmean(X, axis=0)
    if not np.allclose(mean_2, 0):
        X -= mean_2
    return X


def log_spectrogram(wav_arr, n_frames, divide_by=None):
    """The frame loop of wav_util.py:66-76 / :101-111: per frame, raw samples
    (NOT the pre-emphasised ones) * Hamming -> scipy.fftpack.fft -> abs -> bins
    0..199 -> log(x + 1)."""
    w = hamming400()
    data_input = np.zeros((n_frames, N_BINS), dtype=np.float64)
    for i in range(0, n_frames):
        p_start = i * FRAME_HOP
        p_end = p_start + FRAME_LEN
        data_line = wav_arr[p_start:p_end]
        data_line = data_line * w
        data_line = np.abs(fft(data_line))
        if divide_by is not None:
            data_line = data_line / divide_by
        data_input[i] = data_line[0:N_BINS]
    return np.log(data_input + 1)


def compute_fbank(wavsignal, fs=16000):
    """wav_util.py:49-79 on a decoded mono signal (int16 or float32)."""
    wav_arr = np.array(wavsignal)
    n = n_frames_fbank(len(wav_arr), fs)
    feature = log_spectrogram(wav_arr, n)
    return zscore_columns(feature)


def compute_fbank_unnormalised(wavsignal, fs=16000):
    """wav_util.py:49-76: ``feature`` just before the z-score of :79."""
    wav_arr = np.array(wavsignal)
    return log_spectrogram(wav_arr, n_frames_fbank(len(wav_arr), fs))


def compute_fbank_from_asrt(wavsignal, fs=16000):
    """wav_util.py:82-112 on a decoded mono signal: one frame fewer, magnitude
    divided by the signal length, no z-score."""
    wav_arr = np.array(wavsignal)
    n = n_frames_asrt(len(wav_arr), fs)
    return log_spectrogram(wav_arr, n, divide_by=len(wav_arr))


def snr2k(signal, noise, dB):
    """noise.py:48-52, with the dtypes numpy 2.x gives float32 inputs (all
    float32; the Python float 10**(-dB/20) is a weak scalar)."""
    energe_s = np.sum(signal * signal) / len(signal)
    energe_n = np.sum(noise * noise) / len(noise)
    K = np.sqrt(energe_s / energe_n) * (10 ** (-dB / 20))
    return K


def mix_noise(signal, noise, dB):
    """noise.py:107-108."""
    K = snr2k(signal, noise, dB)
    return (signal + K * noise).astype(np.float32)


def color_noise_from_normal(x_random, type_noise):
    """noise.py:17-34 with the N(0,1) draw passed in (the reference uses the global numpy RNG at
    :18; parity is defined given the same draw).  Same numpy operations in the same order (the
    result is compared bit for bit with the golden vectors): full FFT, bins 0..floor(N/2) times
    (k+1)^colour (:19-23), the upper bins rebuilt as conjugates of the lower ones -- without the
    Nyquist bin when N is even (:24-27) -- real part of the inverse FFT, minus the mean, divided by
    the maximum (:28-31), float32 (:33)."""
    N = len(x_random)
    keep = int(np.ceil((N + 1) / 2))
    lower = np.fft.fft(x_random)[:keep] * (np.arange(1, keep + 1) ** type_noise)
    mirror = lower[-2:0:-1] if N % 2 == 0 else lower[-1:0:-1]
    y = np.real(np.fft.ifft(np.concatenate([lower, np.conj(mirror)])))
    y = y - np.mean(y)
    return (y / np.max(y)).astype(np.float32)
