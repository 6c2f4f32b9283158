"""GPU: pipeline.HostRoundTrip (the end-to-end call bench.py times: pinned host buffers in, pinned host buffers out,
double-buffered) returns exactly what the device-resident step returns -- loss, features and gradient bit for bit,
gradient padding rows untouched (zero) -- for both ways of bringing the logits in (DMA of the padded tensor,
zero-copy staging of the valid rows), over more submissions than there are slots."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

pytestmark = pytest.mark.gpu


def _batch(seed, B=24):
    from oracle import synth
    rng = np.random.default_rng(seed)
    lens = synth.ragged_lengths(rng, B, 1.0, 3.0)
    pcm = [synth.g2_voiced(rng, int(n)) for n in lens]
    il = np.array([synth.t_ctc(synth.n_frames(int(n))) for n in lens], dtype=np.int32)
    x, labels, ll, il = synth.ctc_batch(rng, il, 64, 2, 6, lmax=8)
    return pcm, x, labels, ll, il


@pytest.mark.parametrize("logits_in", ["dma", "zero_copy"])
def test_round_trip_equals_device_step(logits_in):
    from asr_dfcnn_transformer_b200 import features, pipeline
    dev = torch.device("cuda", 0)
    batches = [_batch(900 + i) for i in range(3)]
    T = max(b[1].shape[0] for b in batches)
    B, V = batches[0][1].shape[1:]
    packs = [features.pack_host(b[0], pin=True) for b in batches]
    step = pipeline.HotPathStep(dev)
    rt = pipeline.HostRoundTrip(step, max(int(p.frame_offsets[-1]) for p in packs), max(p.samples.numel() for p in packs),
                                packs[0].samples.dtype, T, B, V, batches[0][2].shape[1], slots=2, logits_in=logits_in)
    want, slots = [], []
    for (pcm, x, labels, ll, il), pk in zip(batches, packs):
        xp = np.zeros((T, B, V), np.float32)
        xp[:x.shape[0]] = x
        h_logits = torch.from_numpy(xp).pin_memory()
        h_labels = torch.from_numpy(labels.astype(np.int32)).pin_memory()
        so = torch.as_tensor(np.asarray(pk.sample_offsets, dtype=np.int64)).to(dev)
        sc = torch.as_tensor(np.asarray(pk.sample_counts, dtype=np.int64)).to(dev)
        fo = torch.as_tensor(np.asarray(pk.frame_offsets, dtype=np.int64)).to(dev)
        dll, dil = torch.as_tensor(ll).to(dev), torch.as_tensor(il).to(dev)
        nf = int(pk.frame_offsets[-1])
        # the device-resident step on the same inputs
        feats, res = step(pk.samples.to(dev), so, sc, fo, B, nf, h_logits.to(dev), h_labels.to(dev), dll, dil, V - 1)
        torch.cuda.synchronize()
        want.append((res.loss.cpu().numpy(), feats.cpu().numpy(), res.grad.cpu().numpy(), il, nf))
        s = rt.submit(pk.samples, so, sc, fo, B, nf, h_logits, h_labels, dll, dil, V - 1, valid_rows=int(il.sum()))
        assert s.d2h_bytes == 4 * B + nf * 800 + int(il.sum()) * V * 4
        assert s.h2d_bytes >= pk.samples.numel() * 2 + int(il.sum()) * V * 4
        slots.append(s)
        if len(slots) >= 2:                 # a slot is reused by the submit after next: read it before that
            k = len(slots) - 2
            _check(rt, slots[k], want[k], T)
    _check(rt, slots[-1], want[-1], T)
    rt.drain()


def _check(rt, s, want, T):
    loss, feats, grad, il, nf = want
    rt.wait(s)
    assert np.array_equal(s.h_loss.numpy(), loss)
    assert np.array_equal(s.h_feat[:nf].numpy(), feats)
    got = s.h_grad.numpy()
    valid = np.arange(T)[:, None] < il[None, :]
    assert np.array_equal(got[valid], grad[valid])
    assert not got[~valid].any()
