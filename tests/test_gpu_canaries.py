"""GPU: no kernel writes outside the buffers it was given.  (compute-sanitizer is not available on
the GPU pool, so the outputs are carved out of larger sentinel-filled allocations and the guard
bands are checked after the call; the workspaces are the library's own and are sized by it.)"""
import numpy as np
import pytest

from oracle import synth

pytestmark = pytest.mark.gpu

SENT = 12345.678


def _guarded(torch, shape, dtype, dev, guard=1024):
    n = int(np.prod(shape))
    big = torch.full((n + 2 * guard,), SENT if dtype.is_floating_point else 123456789, dtype=dtype, device=dev)
    return big, big[guard:guard + n].view(*shape), guard


def _intact(big, guard, n):
    ref = big[0].item()
    return bool((big[:guard] == ref).all()) and bool((big[guard + n:] == ref).all())


def test_spectrogram_stays_inside_its_output():
    import torch
    from asr_dfcnn_transformer_b200 import features
    rng = np.random.default_rng(3)
    sigs = [synth.g2_voiced(rng, n) for n in (16000, 403, 48123, 7777, 160 * 31 + 400)]
    pk = features.pack_host(sigs, 16000, "fbank")
    dev = torch.device("cuda", 0)
    big, out, g = _guarded(torch, (pk.total_frames, 200), torch.float32, dev)
    features.spectrogram_device(pk.samples.to(dev), torch.from_numpy(pk.sample_offsets).to(dev),
                                torch.from_numpy(pk.sample_counts).to(dev), torch.from_numpy(pk.frame_offsets).to(dev),
                                len(sigs), pk.total_frames, "fbank", out=out)
    torch.cuda.synchronize()
    assert _intact(big, g, out.numel()) and bool(torch.isfinite(out).all())


def test_ctc_stays_inside_its_outputs():
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    rng = np.random.default_rng(4)
    dev = torch.device("cuda", 0)
    for il, lab_hi in ((rng.integers(5, 60, 9).astype(np.int32), 20),        # fused kernel
                       (np.array([300, 120, 7], np.int32), 60)):             # generic kernels (long lattices)
        x, labels, ll, il = synth.ctc_batch(rng, il, 1424, 2, lab_hi, lmax=64)
        T, B, V = x.shape
        gbig, grad, g = _guarded(torch, (T, B, V), torch.float32, dev)
        lbig, loss, _ = _guarded(torch, (B,), torch.float32, dev)
        sbig, status, _ = _guarded(torch, (B,), torch.int32, dev)
        tbig, tokens, _ = _guarded(torch, (B, T), torch.int32, dev)
        nbig, tlen, _ = _guarded(torch, (B,), torch.int32, dev)
        qbig, nsl, _ = _guarded(torch, (B,), torch.float32, dev)
        r = ctc.ctc_loss_grad(torch.as_tensor(x).to(dev), labels, ll, il, V - 1, decode=True, grad_out=grad,
                              outputs=(loss, grad, status, tokens, tlen, nsl))
        torch.cuda.synchronize()
        assert r.grad.data_ptr() == grad.data_ptr() and r.loss.data_ptr() == loss.data_ptr()
        for big, view in ((gbig, grad), (lbig, loss), (sbig, status), (tbig, tokens), (nbig, tlen), (qbig, nsl)):
            assert _intact(big, g, view.numel())
        assert int(status.max()) == 0 and bool(torch.isfinite(loss).all())


def test_mel_front_end_stays_inside_its_output():
    import torch
    from asr_dfcnn_transformer_b200 import _lib, wav_util
    rng = np.random.default_rng(5)
    dev = torch.device("cuda", 0)
    sigs = [rng.standard_normal(n) * 0.1 for n in (16000, 401, 399, 5173, 160 * 7 + 400)]
    counts = np.array([len(s) for s in sigs], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    nfr = np.array([wav_util.logfbank_frames(int(n)) for n in counts], dtype=np.int64)
    fo = np.concatenate([[0], np.cumsum(nfr)]).astype(np.int64)
    x_d = torch.from_numpy(np.concatenate(sigs)).to(dev)
    so, sc, fo_d = (torch.from_numpy(a).to(dev) for a in (offs[:-1].copy(), counts, fo))
    bins = torch.from_numpy(wav_util._mel_bins(200, 512, 16000)).to(dev)
    big, out, g = _guarded(torch, (int(fo[-1]), 200), torch.float32, dev)
    for normalise in (0, 1):
        st = _lib.lib().asrk_logfbank_run(_lib.ptr(x_d), _lib.ptr(so), _lib.ptr(sc), _lib.ptr(fo_d), None, _lib.ptr(bins),
                                          len(sigs), int(fo[-1]), 200, 400, 160, 0.97, normalise, _lib.ptr(out),
                                          _lib.stream_ptr(None))
        assert st == 0
        torch.cuda.synchronize()
        assert _intact(big, g, out.numel()) and bool(torch.isfinite(out).all())
