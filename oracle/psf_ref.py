"""ORACLE (test infrastructure only): the mel filterbank front end the reference's live
loaders call -- ``compute_fbank_from_api`` (/root/reference/util/wav_util.py:22-31) =
``python_speech_features.logfbank(signal, sample_rate, nfilt=200)`` followed by
``sklearn.preprocessing.scale``.

python_speech_features is a third-party dependency that is NOT vendored in the reference and
not installable here (pin: ``python-speech-features==0.6``, requirements.txt:42), so this is
a restatement of its published algorithm (base.py: fbank / logfbank / get_filterbanks /
hz2mel / mel2hz; sigproc.py: preemphasis / framesig / magspec / powspec), float64 numpy like
the original.  PARITY UNPINNED by the reference (it has no tests / golden vectors); anchored
on the call site (defaults winlen=0.025, winstep=0.01, nfft=512, lowfreq=0, highfreq=fs/2,
preemph=0.97, rectangular window) and on structural known answers in tests/test_oracle_psf.py.
"""
import decimal
import math

import numpy as np


def round_half_up(number):
    return int(decimal.Decimal(number).quantize(decimal.Decimal("1"), rounding=decimal.ROUND_HALF_UP))


def hz2mel(hz):
    return 2595 * np.log10(1 + hz / 700.0)


def mel2hz(mel):
    return 700 * (10 ** (mel / 2595.0) - 1)


def preemphasis(signal, coeff=0.97):
    return np.append(signal[0], signal[1:] - coeff * signal[:-1])


def num_frames(slen, frame_len=400, frame_step=160):
    if slen <= frame_len:
        return 1
    return 1 + int(math.ceil((1.0 * slen - frame_len) / frame_step))


def framesig(sig, frame_len, frame_step):
    slen = len(sig)
    frame_len = int(round_half_up(frame_len))
    frame_step = int(round_half_up(frame_step))
    numframes = num_frames(slen, frame_len, frame_step)
    padlen = int((numframes - 1) * frame_step + frame_len)
    padsignal = np.concatenate((sig, np.zeros((padlen - slen,))))
    idx = np.arange(frame_len)[None, :] + (np.arange(numframes) * frame_step)[:, None]
    return padsignal[idx]          # winfunc = ones


def powspec(frames, nfft):
    return 1.0 / nfft * np.square(np.absolute(np.fft.rfft(frames, nfft)))


def mel_bins(nfilt=200, nfft=512, samplerate=16000, lowfreq=0, highfreq=None):
    highfreq = highfreq or samplerate / 2
    melpoints = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    return np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)


def get_filterbanks(nfilt=200, nfft=512, samplerate=16000, lowfreq=0, highfreq=None):
    bin = mel_bins(nfilt, nfft, samplerate, lowfreq, highfreq)
    fbank = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(0, nfilt):
        for i in range(int(bin[j]), int(bin[j + 1])):
            fbank[j, i] = (i - bin[j]) / (bin[j + 1] - bin[j])
        for i in range(int(bin[j + 1]), int(bin[j + 2])):
            fbank[j, i] = (bin[j + 2] - i) / (bin[j + 2] - bin[j + 1])
    return fbank


def logfbank(signal, samplerate=16000, nfilt=200, nfft=512, preemph=0.97):
    signal = preemphasis(np.asarray(signal, dtype=np.float64), preemph)
    frames = framesig(signal, 0.025 * samplerate, 0.01 * samplerate)
    pspec = powspec(frames, nfft)
    fb = get_filterbanks(nfilt, nfft, samplerate, 0, samplerate / 2)
    feat = np.dot(pspec, fb.T)
    feat = np.where(feat == 0, np.finfo(float).eps, feat)
    return np.log(feat)


def scale(x):
    """sklearn.preprocessing.scale on columns (as oracle/fbank_ref.py)."""
    mean = x.mean(axis=0)
    std = x.std(axis=0)
    std = np.where(std < 10 * np.finfo(np.float64).eps, 1.0, std)
    return (x - mean) / std


def compute_fbank_from_api(signal, sample_rate=16000, nfilt=200):
    """wav_util.py:22-31."""
    return scale(logfbank(signal, sample_rate, nfilt=nfilt))
