"""Batched spectrogram features on the GPU (host side of asrk_spectrogram_run).

The per-file functions of the reference (``util/wav_util.py``) are mirrored in
``asr_dfcnn_transformer_b200/wav_util.py``; this module is the batched addition:
a ragged batch of utterances goes through ONE launch sequence.

Host rules copied from the reference (they are control logic, not arithmetic):
  * number of frames = ``int(N / fs * 1000 - 25) // 10 + 1`` evaluated as that
    very Python float expression (wav_util.py:61; asrt variant without the +1,
    wav_util.py:96) -- it differs from integer arithmetic for some N.
"""
import ctypes
from collections import namedtuple

import numpy as np

from . import _lib

FRAME_LEN, FRAME_HOP, N_BINS = 400, 160, 200
MODES = {"fbank": _lib.SPEC_FBANK, "asrt": _lib.SPEC_ASRT, "fbank_raw": _lib.SPEC_FBANK_RAW}

PackedBatch = namedtuple("PackedBatch", "samples noise sample_offsets sample_counts frame_offsets "
                                        "n_frames total_frames")
FeatureBatch = namedtuple("FeatureBatch", "features frame_offsets n_frames")


def n_frames_for(n_samples, fs=16000, mode="fbank"):
    """wav_util.py:61 / :96 -- keep the float expression as is."""
    n = int(n_samples / fs * 1000 - 25) // 10
    if mode != "asrt":
        n += 1
    return n


def _check_frames(n, n_samples):
    if n <= 0:
        return 0
    if FRAME_HOP * (n - 1) + FRAME_LEN > n_samples:
        # the reference would fail on ``data_line * w`` with a short last frame
        # (sample rates below 16 kHz, wav_util.py:67-71)
        raise ValueError("frame count %d needs more samples than the utterance has (%d); "
                         "the reference assumes 16 kHz audio" % (n, n_samples))
    return n


_workspaces = {}


def workspace(nbytes, device, tag="spec", stream=None):
    """grow-only scratch tensor per (device, tag, LAUNCH stream): the stream the kernels are
    enqueued on (``stream``; torch's current stream when None) owns the buffer, so two calls on
    two explicit streams never share scratch memory, and a grown buffer is allocated -- and its
    predecessor released -- under the stream that uses it."""
    torch = _lib.require_cuda()
    s = torch.cuda.current_stream(device) if stream is None else stream
    key = (str(device), tag, s.cuda_stream)
    w = _workspaces.get(key)
    if w is None or w.numel() < nbytes:
        with torch.cuda.stream(s):
            w = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = w
    return w


def pack_host(signals, fs=16000, mode="fbank", noises=None, pin=True):
    """Pack a list of 1-D int16 / float32 arrays into one ragged (pinned) host
    buffer whose utterance starts are 16-byte aligned.  Returns CPU tensors /
    numpy offset arrays."""
    import torch
    if len(signals) == 0:
        raise ValueError("empty batch")
    first = np.asarray(signals[0])
    if first.dtype == np.int16:
        dtype, tdtype, align = np.int16, torch.int16, 8
    else:
        dtype, tdtype, align = np.float32, torch.float32, 4
    B = len(signals)
    counts = np.zeros(B, dtype=np.int64)
    offs = np.zeros(B, dtype=np.int64)
    nfr = np.zeros(B, dtype=np.int64)
    pos = 0
    for i, s in enumerate(signals):
        a = np.asarray(s)
        if a.ndim != 1:
            raise ValueError("signals must be mono 1-D arrays")
        n = int(a.shape[0])
        counts[i] = n
        offs[i] = pos
        nfr[i] = _check_frames(n_frames_for(n, fs, mode), n)
        pos += (n + align - 1) // align * align
    if noises is not None and counts.max() > (1 << 22):
        # asrk_snr2k_run's summation tree (include/asrk.h): a longer utterance would get a NaN gain
        raise ValueError("noise mix: utterances of more than 2**22 samples are not supported (device SNR2K tree)")
    total = max(pos, align)
    pin = pin and torch.cuda.is_available()
    buf = torch.zeros(total, dtype=tdtype, pin_memory=pin)
    nbuf = torch.zeros(total, dtype=torch.float32, pin_memory=pin) if noises is not None else None
    bnp = buf.numpy()
    for i, s in enumerate(signals):
        a = np.asarray(s)
        bnp[offs[i]: offs[i] + counts[i]] = a.astype(dtype, copy=False)
        if nbuf is not None:
            nn = np.asarray(noises[i], dtype=np.float32)
            if nn.shape[0] != counts[i]:
                raise ValueError("noise length must equal the signal length (noise.py:106)")
            nbuf.numpy()[offs[i]: offs[i] + counts[i]] = nn
    frame_offsets = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(nfr, out=frame_offsets[1:])
    return PackedBatch(buf, nbuf, offs, counts, frame_offsets, nfr, int(frame_offsets[-1]))


def spectrogram_device(samples, sample_offsets, sample_counts, frame_offsets, batch, total_frames,
                       mode="fbank", noise=None, gain=None, snr_db=None, out=None,
                       out_row_offsets=None, stream=None, phases=_lib.PHASE_ALL, cta_limit=0):
    """Launch the kernels on device tensors that are already packed.

    samples: int16 / float32 [total]; sample_offsets, sample_counts: int64 [B];
    frame_offsets: int64 [B+1] (all device); out: float32 [rows, 200] (allocated
    if None).  No synchronisation."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    dev = samples.device
    if samples.dtype == torch.int16:
        dt = _lib.DTYPE_I16
    elif samples.dtype == torch.float32:
        dt = _lib.DTYPE_F32
    else:
        raise TypeError("samples must be int16 or float32")
    if out is None:
        out = torch.empty((max(total_frames, 1), N_BINS), dtype=torch.float32, device=dev)[:total_frames]
    nbytes = L.asrk_spectrogram_workspace_bytes(batch, total_frames)
    ws = workspace(nbytes, dev, "spec", stream)
    st = L.asrk_spectrogram_run_phases(_lib.ptr(samples), dt, _lib.ptr(noise), _lib.ptr(gain),
                                       _lib.ptr(snr_db), _lib.ptr(sample_offsets), _lib.ptr(sample_counts),
                                       _lib.ptr(frame_offsets), _lib.ptr(out_row_offsets), batch,
                                       total_frames, MODES[mode], _lib.ptr(out), _lib.ptr(ws), ws.numel(),
                                       _lib.stream_ptr(stream), int(phases) | (int(cta_limit) << 16))
    _lib.check(st, "asrk_spectrogram_run")
    return out


ZScoreWork = namedtuple("ZScoreWork", "features stats ticket frame_offsets row_offsets batch total_frames")


def zscore_work(features, frame_offsets, batch, total_frames, out_row_offsets=None, stream=None):
    """The z-score pass of a ``spectrogram_device(..., phases=SETUP | MAIN | STATS)`` call on ``stream``, packaged
    for ``ctc.ctc_loss_grad(zscore=...)``: the fused CTC kernel normalises the rows as co-work
    (``asrk_ctc_loss_grad_zscore_run``).  ``stats`` / ``ticket`` are addresses inside that call's workspace."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    ws = workspace(L.asrk_spectrogram_workspace_bytes(batch, total_frames), features.device, "spec", stream)
    stats, ticket = ctypes.c_void_p(), ctypes.c_void_p()
    rc = L.asrk_spectrogram_zscore_handles(_lib.ptr(ws), ws.numel(), int(batch), int(total_frames),
                                           ctypes.byref(stats), ctypes.byref(ticket))
    _lib.check(rc, "asrk_spectrogram_zscore_handles")
    return ZScoreWork(features, stats, ticket, frame_offsets, out_row_offsets, int(batch), int(total_frames))


def compute_features(signals, fs=16000, mode="fbank", noises=None, snr_db=None, gain=None,
                     device=None, padded_rows=None):
    """Features of a list of utterances (host arrays) in one batched GPU pass.

    Returns FeatureBatch(features float32 device tensor, frame_offsets, n_frames):
    ragged ``[total_frames, 200]`` by default, or zero-padded
    ``[B, padded_rows, 200]`` (the loader layout of data_loader.py:107,146)."""
    torch = _lib.require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    pk = pack_host(signals, fs, mode, noises)
    B = len(signals)
    samples = pk.samples.to(dev, non_blocking=True)
    noise = pk.noise.to(dev, non_blocking=True) if pk.noise is not None else None
    so = torch.from_numpy(pk.sample_offsets).to(dev, non_blocking=True)
    sc = torch.from_numpy(pk.sample_counts).to(dev, non_blocking=True)
    fo = torch.from_numpy(pk.frame_offsets).to(dev, non_blocking=True)
    g = s = None
    if noises is not None:
        if gain is not None:
            g = torch.as_tensor(np.asarray(gain, dtype=np.float32)).to(dev)
        elif snr_db is not None:
            s = torch.as_tensor(np.asarray(snr_db, dtype=np.int32)).to(dev)
        else:
            raise ValueError("noise mixing needs snr_db or gain")
    out = oro = None
    if padded_rows is not None:
        if int(pk.n_frames.max()) > padded_rows:
            raise ValueError("utterance longer than padded_rows frames (data_loader.py:139-140)")
        out = torch.zeros((B, padded_rows, N_BINS), dtype=torch.float32, device=dev)
        oro = (torch.arange(B, dtype=torch.int64) * padded_rows).to(dev)
    res = spectrogram_device(samples, so, sc, fo, B, pk.total_frames, mode, noise, g, s,
                             out=out.view(-1, N_BINS) if out is not None else None, out_row_offsets=oro)
    return FeatureBatch(out if out is not None else res, pk.frame_offsets, pk.n_frames)
