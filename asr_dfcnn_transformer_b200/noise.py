"""Drop-in for the mixing half of the reference's ``util/noise.py``.

``SNR2K`` (noise.py:48-52) and the mix of noise.py:108 run on the GPU; the mix is
normally not materialised at all but fused into the feature kernel's load stage
(``features.compute_features(..., noises=..., snr_db=...)``).  ``color_noise``
(noise.py:17-34) draws its normal deviates on the host exactly like the reference
(numpy's global generator, so ``np.random.seed`` reproduces the reference's noise) and
runs the two arbitrary-length FFTs, the spectral shaping and the normalisation on the device.
"""
import numpy as np

from . import _lib
from . import features


def color_noise_batch(normals, colours, device=None):
    """Coloured noise for a batch: ``normals`` is a list of float64 arrays of N(0,1) deviates
    (one per utterance, any lengths), ``colours`` the exponents in [-1, 1].  Returns
    (float32 device tensor with the noises back to back, int64 host offsets [B+1])."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    xs = [np.ascontiguousarray(np.asarray(x, dtype=np.float64)) for x in normals]
    if len(xs) == 0 or any(x.ndim != 1 or x.shape[0] == 0 for x in xs):
        raise ValueError("normals must be non-empty 1-D arrays")
    if len(colours) != len(xs):
        raise ValueError("one colour per utterance")
    counts = np.array([len(x) for x in xs], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    B, nmax = len(xs), int(counts.max())
    L = _lib.lib()
    x_d = torch.from_numpy(np.concatenate(xs)).to(dev)
    o_d = torch.from_numpy(offs[:-1].copy()).to(dev)
    c_d = torch.from_numpy(counts).to(dev)
    col_d = torch.from_numpy(np.asarray(colours, dtype=np.float64)).to(dev)
    out = torch.empty(int(offs[-1]), dtype=torch.float32, device=dev)
    nbytes = L.asrk_color_noise_workspace_bytes(B, nmax)
    ws = features.workspace(nbytes, dev, "cnoise")
    st = L.asrk_color_noise_run(_lib.ptr(x_d), _lib.ptr(o_d), _lib.ptr(c_d), _lib.ptr(col_d), B, nmax,
                                _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.stream_ptr())
    _lib.check(st, "asrk_color_noise_run")
    return out, offs


def color_noise(len_noise, type_noise):
    """noise.py:17-34: one coloured noise of ``len_noise`` samples, float32, zero mean,
    maximum 1.  The deviates come from ``np.random.normal(0, 1, len_noise)`` like in the
    reference, so the same ``np.random.seed`` gives the same noise."""
    x_random = np.random.normal(0, 1, len_noise)
    out, _ = color_noise_batch([x_random], [float(type_noise)])
    return out.cpu().numpy()


def _pack_pair(signal, noise, dev, torch):
    s = torch.as_tensor(np.ascontiguousarray(signal, dtype=np.float32)).to(dev)
    n = torch.as_tensor(np.ascontiguousarray(noise, dtype=np.float32)).to(dev)
    return s, n


SNR2K_MAX_SAMPLES = 1 << 22


def SNR2K(signal, noise, dB):
    """noise.py:48-52 -> numpy float32 scalar (the dtype numpy 2.x gives for
    float32 inputs), computed on the device with numpy's summation order.
    One limit the reference does not have: the device tree of numpy's pairwise
    sum holds up to 2**22 samples (262 s at 16 kHz); longer inputs raise."""
    if np.size(signal) > SNR2K_MAX_SAMPLES:
        raise ValueError("SNR2K: %d samples; the device summation tree holds up to 2**22 (include/asrk.h, "
                         "asrk_snr2k_run)" % np.size(signal))
    torch = _lib.require_cuda()
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    s, n = _pack_pair(signal, noise, dev, torch)
    if s.numel() != n.numel():
        raise ValueError("operands could not be broadcast together")
    offs = torch.zeros(1, dtype=torch.int64, device=dev)
    cnt = torch.full((1,), s.numel(), dtype=torch.int64, device=dev)
    db = torch.full((1,), int(dB), dtype=torch.int32, device=dev)
    out = torch.empty(1, dtype=torch.float32, device=dev)
    rc = L.asrk_snr2k_run(_lib.ptr(s), _lib.ptr(n), _lib.ptr(offs), _lib.ptr(cnt), _lib.ptr(db), 1,
                          _lib.ptr(out), _lib.stream_ptr())
    _lib.check(rc, "asrk_snr2k_run")
    return np.float32(out.cpu().numpy()[0])


def mix(signal, noise, dB):
    """noise.py:107-108: ``(signal + K * noise).astype(np.float32)`` (host result;
    used by tests and by add_noise's in-memory branch)."""
    torch = _lib.require_cuda()
    K = SNR2K(signal, noise, dB)
    dev = torch.device("cuda", torch.cuda.current_device())
    s, n = _pack_pair(signal, noise, dev, torch)
    return (s + torch.tensor(K, dtype=torch.float32, device=dev) * n).cpu().numpy()


def _load_signal(item, sample_rate):
    """librosa.load(file, sr=sample_rate) for the cases that need no resampling: 16-bit / float
    PCM wavs at ``sample_rate`` -> mono float32 in [-1, 1].  Arrays pass through."""
    if isinstance(item, str):
        import scipy.io.wavfile as wavfile
        fs, sig = wavfile.read(item)
        if fs != sample_rate:
            raise ValueError("resampling is not part of this path: %s is %d Hz" % (item, fs))
        sig = np.asarray(sig)
        if sig.ndim == 2:
            sig = sig.mean(axis=1)
        if sig.dtype == np.int16:
            return (sig.astype(np.float32) / 32768.0).astype(np.float32)
        return sig.astype(np.float32)
    return np.ascontiguousarray(item, dtype=np.float32)


def add_noise(signal_path, n_to_add=1, sample_rate=16000, out_path=None, dB="random", type_noise="random",
              keep_bits=False):
    """noise.py:70-128.  ``signal_path``: a list of wav files (or of decoded float32 signals), or a
    directory.  For every signal and every copy: SNR ~ randint(5, 10) and colour ~ randint(-10, 10)/10
    from ``random`` (same draws, same order as the reference), coloured noise of the signal's length,
    gain from the SNR, ``(signal + K * noise).astype(float32)`` -- noise, gain and mix on the device.
    Returns (list of mixed signals, list of written file names); bad arguments print and return 0
    like the reference."""
    import os
    import random
    if isinstance(signal_path, (list, tuple)):
        if len(signal_path) == 0 or (isinstance(signal_path[0], str) and not os.path.isfile(signal_path[0])):
            print("Error signal_path!")
            return 0
        items = list(signal_path)
    elif isinstance(signal_path, str) and os.path.isdir(signal_path):
        items = [os.path.join(signal_path, f) for f in os.listdir(signal_path)]
    else:
        print("Error signal_path!")
        return 0
    out, names = [], []
    for l, item in enumerate(items):
        signal = _load_signal(item, sample_rate)
        for n in range(n_to_add):
            snr_dB = random.randint(5, 10) if dB == "random" else int(dB)
            if type_noise == "random":
                type_n = random.randint(-10, 10) / 10
            else:
                type_n = float(type_noise)
                if abs(type_n) > 1:
                    print("Error noise type! Please given a float belongs to [-1, 1] !")
                    return 0
            nz = color_noise(len(signal), type_n)
            mixed = mix(signal, nz, snr_dB)
            if out_path is not None:
                import scipy.io.wavfile as wavfile
                path = out_path + "/" + str(l) + "_" + str(n) + "_" + str(type_n) + "_" + str(snr_dB) + "_dB.wav"
                names.append(path)
                y = mixed / np.abs(mixed).max() if np.abs(mixed).max() > 1 else mixed     # noise.py:115-119
                wavfile.write(path, sample_rate, y.astype(np.float32))
            else:
                out.append(mixed)
    return out, names


def noisy_features(signals, noises, snr_db, **kw):
    """Features of the noise-mixed signals without materialising the mix."""
    return features.compute_features(signals, noises=noises, snr_db=snr_db, **kw)
