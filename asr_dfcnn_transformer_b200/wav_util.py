"""Drop-in for the reference's ``util/wav_util.py`` feature functions.

Same names, arguments and return types (float64 numpy arrays) as
/root/reference/util/wav_util.py:13-112, computed by the CUDA kernels behind
include/asrk.h.  ``compute_fbank_batch`` is the batched addition.
"""
import wave

import numpy as np

from . import features


def read_wav_data(filename):
    """wav_util.py:34-45 -- returns (int16 [channels, samples], framerate)."""
    wav = wave.open(filename, "rb")
    num_frame = wav.getnframes()
    num_channel = wav.getnchannels()
    framerate = wav.getframerate()
    str_data = wav.readframes(num_frame)
    wav.close()
    wave_data = np.frombuffer(str_data, dtype=np.short).copy()
    wave_data.shape = -1, num_channel
    return wave_data.T, framerate


def _read_mono(file):
    """scipy.io.wavfile.read as used at wav_util.py:53 (int16 or float32 PCM)."""
    import scipy.io.wavfile as wav
    fs, sig = wav.read(file)
    sig = np.asarray(sig)
    if sig.ndim != 1:
        raise ValueError("compute_fbank expects a mono wav (wav_util.py:66-69)")
    if sig.dtype != np.int16:
        sig = sig.astype(np.float32)
    return fs, sig


def compute_fbank(file):
    """wav_util.py:49-79: Hamming(400) frames at hop 160, |FFT| bins 0..199,
    log(x+1), per-column z-score.  Returns float64 ``[n_frames, 200]``."""
    fs, sig = _read_mono(file)
    fb = features.compute_features([sig], fs=fs, mode="fbank")
    return fb.features.cpu().numpy().astype(np.float64)


def compute_fbank_from_asrt(file):
    """wav_util.py:82-112: one frame fewer, magnitude / signal length, no z-score."""
    wavsignal, fs = read_wav_data(file)
    n_total = wavsignal.shape[1]
    # wav_util.py:94 takes len() of the [C,N] array's first row after np.append on
    # row 0, i.e. the mono length; channel 0 is the one transformed (:104)
    sig = np.ascontiguousarray(wavsignal[0])
    fb = features.compute_features([sig], fs=fs, mode="asrt")
    assert n_total == sig.shape[0]
    return fb.features.cpu().numpy().astype(np.float64)


def compute_fbank_batch(signals, fs=16000, mode="fbank", **kw):
    """Batched GPU addition: list of int16 / float32 1-D arrays -> FeatureBatch with
    a float32 device tensor (ragged ``[sum n_frames, 200]`` or padded)."""
    return features.compute_features(signals, fs=fs, mode=mode, **kw)


def compute_fbank_from_api(signal, sample_rate, nfilt=200):
    """wav_util.py:22-31 (``python_speech_features.logfbank`` + scale): the mel
    filterbank front end is the next row of the scope table and is
    not part of this round's hot path."""
    raise NotImplementedError("logfbank front end (SURVEY.md section 8f row 2) is not built yet; "
                              "use compute_fbank / compute_fbank_batch")


def compute_fbank_from_file(file, feature_dim=200, sf_flag=False):
    """wav_util.py:13-19."""
    raise NotImplementedError("logfbank front end (SURVEY.md section 8f row 2) is not built yet; "
                              "use compute_fbank / compute_fbank_batch")
