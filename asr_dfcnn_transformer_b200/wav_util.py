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


def _mel_bins(nfilt, nfft, samplerate):
    """FFT bin edges of the triangular mel filters: python_speech_features get_filterbanks
    (lowfreq=0, highfreq=fs/2), evaluated on the host with the library's own numpy expression
    so that the floor() lands on the same bins."""
    def hz2mel(hz):
        return 2595 * np.log10(1 + hz / 700.0)

    def mel2hz(mel):
        return 700 * (10 ** (mel / 2595.0) - 1)
    melpoints = np.linspace(hz2mel(0), hz2mel(samplerate / 2), nfilt + 2)
    return np.floor((nfft + 1) * mel2hz(melpoints) / samplerate).astype(np.int32)


def _round_half_up(number):
    import decimal
    return int(decimal.Decimal(number).quantize(decimal.Decimal("1"), rounding=decimal.ROUND_HALF_UP))


def logfbank_frames(n_samples, frame_len=400, frame_step=160):
    """sigproc.framesig: 1 frame up to frame_len samples, else 1 + ceil((N - len) / step)."""
    import math
    if n_samples <= frame_len:
        return 1
    return 1 + int(math.ceil((1.0 * n_samples - frame_len) / frame_step))


def compute_fbank_from_api_batch(signals, sample_rate=16000, nfilt=200, normalise=True, padded_rows=None,
                                 device=None, preemph=0.97):
    """Batched mel front end on the device: list of 1-D signals (any real dtype, e.g. the
    float64 in [-1, 1] that soundfile returns) -> FeatureBatch with a float32 device tensor,
    ragged ``[sum n_frames, nfilt]`` or zero-padded ``[B, padded_rows, nfilt]``."""
    from . import _lib
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if len(signals) == 0:
        raise ValueError("empty batch")
    frame_len = _round_half_up(0.025 * sample_rate)        # sigproc.round_half_up
    frame_step = _round_half_up(0.01 * sample_rate)
    if frame_len > 512:
        raise ValueError("frame longer than the 512-point transform (sample rate > 20.48 kHz)")
    sigs = [np.ascontiguousarray(np.asarray(s, dtype=np.float64)) for s in signals]
    for s in sigs:
        if s.ndim != 1 or s.shape[0] == 0:
            raise ValueError("signals must be non-empty 1-D arrays")
    counts = np.array([len(s) for s in sigs], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    nfr = np.array([logfbank_frames(int(n), frame_len, frame_step) for n in counts], dtype=np.int64)
    fo = np.concatenate([[0], np.cumsum(nfr)]).astype(np.int64)
    B = len(sigs)
    samples = torch.from_numpy(np.concatenate(sigs)).to(dev)
    so, sc, fo_d = (torch.from_numpy(a).to(dev) for a in (offs[:-1].copy(), counts, fo))
    bins = torch.from_numpy(_mel_bins(nfilt, 512, sample_rate)).to(dev)
    oro = None
    if padded_rows is not None:
        if int(nfr.max()) > padded_rows:
            raise ValueError("utterance longer than padded_rows frames (data_loader.py:139-140)")
        out = torch.zeros((B, padded_rows, nfilt), dtype=torch.float32, device=dev)
        oro = (torch.arange(B, dtype=torch.int64) * padded_rows).to(dev)
    else:
        out = torch.empty((int(fo[-1]), nfilt), dtype=torch.float32, device=dev)
    st = _lib.lib().asrk_logfbank_run(_lib.ptr(samples), _lib.ptr(so), _lib.ptr(sc), _lib.ptr(fo_d), _lib.ptr(oro),
                                      _lib.ptr(bins), B, int(fo[-1]), int(nfilt), frame_len, frame_step, float(preemph),
                                      1 if normalise else 0, _lib.ptr(out), _lib.stream_ptr(None))
    _lib.check(st, "asrk_logfbank_run")
    return features.FeatureBatch(out, fo, nfr)


def compute_fbank_from_api(signal, sample_rate, nfilt=200):
    """wav_util.py:22-31: ``logfbank(signal, sample_rate, nfilt)`` + per-filter z-score.
    Returns float64 ``[n_frames, nfilt]`` like the reference."""
    sig = np.asarray(signal)
    preemph = 0.97
    if sig.ndim == 2:
        # read_wav_data returns [channels, samples] (compute_fbank_from_file without sf_flag, :13-19), and
        # python_speech_features' preemphasis is ``np.append(signal[0], signal[1:] - 0.97 * signal[:-1])``:
        # on a 2-D array that is row 0 UNCHANGED followed by the row differences -- for a mono file the
        # samples as they are, with no pre-emphasis at all.  Reproduced here.
        sig = np.append(sig[0], sig[1:] - 0.97 * sig[:-1])
        preemph = 0.0
    fb = compute_fbank_from_api_batch([sig], sample_rate, nfilt, preemph=preemph)
    return fb.features.cpu().numpy().astype(np.float64)


def compute_fbank_from_file(file, feature_dim=200, sf_flag=False):
    """wav_util.py:13-19.  ``sf_flag`` asks for soundfile (float64 in [-1, 1]); soundfile is
    optional here: 16-bit PCM wavs are read with the wave module and scaled by 1/32768, which
    is exactly what soundfile returns for them."""
    if sf_flag:
        try:
            import soundfile as sf
            signal, sample_rate = sf.read(file)
        except ImportError:
            wavsignal, sample_rate = read_wav_data(file)
            signal = wavsignal[0].astype(np.float64) / 32768.0
    else:
        signal, sample_rate = read_wav_data(file)
    return compute_fbank_from_api(signal, sample_rate, nfilt=feature_dim)
