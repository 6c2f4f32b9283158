"""Drop-in for the mixing half of the reference's ``util/noise.py``.

``SNR2K`` (noise.py:48-52) and the mix of noise.py:108 run on the GPU; the mix is
normally not materialised at all but fused into the feature kernel's load stage
(``features.compute_features(..., noises=..., snr_db=...)``).  ``color_noise``
(noise.py:17-34) generation is the next row of the scope table:
until it has its own kernel the noise is an INPUT of this module.
"""
import numpy as np

from . import _lib
from . import features


def _pack_pair(signal, noise, dev, torch):
    s = torch.as_tensor(np.ascontiguousarray(signal, dtype=np.float32)).to(dev)
    n = torch.as_tensor(np.ascontiguousarray(noise, dtype=np.float32)).to(dev)
    return s, n


def SNR2K(signal, noise, dB):
    """noise.py:48-52 -> numpy float32 scalar (the dtype numpy 2.x gives for
    float32 inputs), computed on the device with numpy's summation order."""
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


def add_noise(signals, noises, dB="random", rng=None):
    """In-memory branch of noise.py:70-128 (``out_path=None``) with the signals
    already decoded and the coloured noise supplied: returns (list of float32
    arrays, name list).  Draws the SNR like noise.py:95-96 when dB == 'random'."""
    import random
    out, names = [], []
    for s, n in zip(signals, noises):
        snr_dB = random.randint(5, 10) if dB == "random" else int(dB)
        out.append(mix(s, n, snr_dB))
    return out, names


def noisy_features(signals, noises, snr_db, **kw):
    """Features of the noise-mixed signals without materialising the mix."""
    return features.compute_features(signals, noises=noises, snr_db=snr_db, **kw)
