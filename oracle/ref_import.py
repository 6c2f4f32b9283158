"""Import the reference's own feature / noise code from /root/reference, unmodified.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Works only in the build
container: /root/reference does not exist on the GPU box, so nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this at run time.  It is
used to (a) validate ``oracle/fbank_ref.py`` and (b) generate the committed
golden vectors (``tools/make_golden.py``).

The reference modules ``util/wav_util.py`` and ``util/noise.py`` do not import
under numpy 2.x / without soundfile, python_speech_features, matplotlib and
librosa.  We install *in-memory* stub modules for those
(none of them is used by the functions we call) and two numpy shims
(``np.float``, binary ``np.fromstring``); the reference source is neither copied
nor edited.
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("ASRK_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "util", "wav_util.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _install_shims():
    def _missing(*_a, **_k):
        raise RuntimeError("stubbed third-party function called in the oracle")

    _stub("soundfile", read=_missing)
    _stub("python_speech_features", logfbank=_missing, mfcc=_missing)
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    _stub("librosa", load=_missing)
    if "tqdm" not in sys.modules:
        try:
            importlib.import_module("tqdm")
        except Exception:  # pragma: no cover
            _stub("tqdm", tqdm=lambda x, *a, **k: x)
    if not hasattr(np, "float"):
        np.float = float  # removed in numpy 1.24; wav_util.py:63-64,98-99
    _orig_fromstring = np.fromstring

    def _fromstring(s, dtype=float, count=-1, sep=""):
        if sep == "" and isinstance(s, (bytes, bytearray, memoryview)):
            return np.frombuffer(s, dtype=dtype, count=count).copy()  # wav_util.py:42
        return _orig_fromstring(s, dtype=dtype, count=count, sep=sep)

    np.fromstring = _fromstring


_cache = {}


def load():
    """Return (wav_util, noise) reference modules."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the reference package is called ``util``; make sure we get that one
    for k in [k for k in sys.modules if k == "util" or k.startswith("util.")]:
        del sys.modules[k]
    noise = importlib.import_module("util.noise")
    wav_util = importlib.import_module("util.wav_util")
    assert os.path.realpath(wav_util.__file__).startswith(os.path.realpath(REFERENCE_ROOT))
    _cache["mods"] = (wav_util, noise)
    return _cache["mods"]


def _wave_sf_read(path):
    """What soundfile.read returns for a 16-bit PCM wav: float64 in [-1, 1] and the sample rate."""
    import wave
    w = wave.open(path, "rb")
    n, ch, fs = w.getnframes(), w.getnchannels(), w.getframerate()
    raw = w.readframes(n)
    w.close()
    a = np.frombuffer(raw, dtype=np.int16).astype(np.float64) / 32768.0
    return (a if ch == 1 else a.reshape(-1, ch)), fs


def load_data_loader():
    """Return the reference's ``lm_and_am.data_loader`` module, unmodified, for generating the loader
    golden vectors (tools/make_golden_loader.py).  keras / tensorflow are stubbed (the loader only imports
    them), ``soundfile.read`` is a wave-module reader, and ``python_speech_features.logfbank`` -- absent
    and un-installable here -- is the restatement in oracle/psf_ref.py: the loader's CONTROL LOGIC (shapes,
    lengths, label ids, reject rules, row deletion) is pinned to the reference's own code, the mel features
    inside remain "parity unpinned"."""
    if "loader" in _cache:
        return _cache["loader"]
    wav_util, _ = load()
    from oracle import psf_ref
    sys.modules["soundfile"].read = _wave_sf_read
    sys.modules["python_speech_features"].logfbank = \
        lambda signal, samplerate=16000, nfilt=26, **kw: psf_ref.logfbank(signal, samplerate, nfilt=nfilt)
    wav_util.sf.read = _wave_sf_read
    wav_util.logfbank = sys.modules["python_speech_features"].logfbank
    keras = _stub("keras")
    ku = _stub("keras.utils", Sequence=object)
    kb = _stub("keras.backend")
    keras.utils, keras.backend = ku, kb
    _stub("tensorflow")
    for k in [k for k in sys.modules if k == "lm_and_am" or k.startswith("lm_and_am.")]:
        del sys.modules[k]
    mod = importlib.import_module("lm_and_am.data_loader")
    assert os.path.realpath(mod.__file__).startswith(os.path.realpath(REFERENCE_ROOT))
    _cache["loader"] = mod
    return mod
