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


@pytest.mark.parametrize("B,kind", [(24, "logits"), (301, "logits"), (5, "prob")])
def test_merged_tail_equals_two_kernel_step(B, kind):
    """pipeline.HotPathStep with the z-score riding on the fused CTC kernel (one kernel for the step's tail) gives
    bit-identical features, losses, gradients and tokens to the separate calls: more utterances than resident CTAs
    (301 > 2 x 148), fewer than the co-work-only CTAs, an utterance without frames, chunk boundaries inside and
    between utterances, both input kinds; and a batch that cannot take the fused kernel falls back silently."""
    from oracle import synth
    from asr_dfcnn_transformer_b200 import _lib, ctc, features, pipeline
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(4200 + B)
    lens = synth.ragged_lengths(rng, B, 0.3, 2.5)
    lens[1] = 300                                    # no frame at all (fbank rule: needs 400 samples)
    pcm = [synth.g2_voiced(rng, int(n)) for n in lens]
    pk = features.pack_host(pcm, pin=False)
    il = np.maximum(1, np.array([synth.t_ctc(max(int(f), 1)) for f in pk.n_frames], dtype=np.int32))
    V = 64
    x, labels, ll, il = synth.ctc_batch(rng, il, V, 1, 4, lmax=8)
    if kind == "prob":
        e = np.exp(x - x.max(-1, keepdims=True))
        x = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    so = torch.as_tensor(np.asarray(pk.sample_offsets, dtype=np.int64)).to(dev)
    sc = torch.as_tensor(np.asarray(pk.sample_counts, dtype=np.int64)).to(dev)
    fo = torch.as_tensor(np.asarray(pk.frame_offsets, dtype=np.int64)).to(dev)
    samples = pk.samples.to(dev)
    nf = int(pk.frame_offsets[-1])
    logits = torch.as_tensor(x).to(dev)
    dlab = torch.as_tensor(labels.astype(np.int32)).to(dev)
    dll, dil = torch.as_tensor(ll).to(dev), torch.as_tensor(il).to(dev)
    bounds = (int(il.max()), int(ll.max()))
    # separate calls
    f_ref = features.spectrogram_device(samples, so, sc, fo, B, nf, "fbank")
    r_ref = ctc.ctc_loss_grad(logits, dlab, dll, dil, V - 1, bounds=bounds, decode=(kind == "logits"), input_kind=kind)
    torch.cuda.synchronize()
    n0 = _lib.lib().asrk_launch_count()
    if kind == "prob":
        # the step itself speaks logits; the co-work entry is reached through the raw op
        f = features.spectrogram_device(samples, so, sc, fo, B, nf, "fbank",
                                        phases=_lib.PHASE_SPEC_SETUP | _lib.PHASE_SPEC_MAIN | _lib.PHASE_SPEC_STATS)
        zw = features.zscore_work(f, fo, B, nf)
        r = ctc.ctc_loss_grad(logits, dlab, dll, dil, V - 1, bounds=bounds, input_kind="prob", zscore=zw)
    else:
        step = pipeline.HotPathStep(dev, merged_tail=True)
        f, r = step(samples, so, sc, fo, B, nf, logits, dlab, dll, dil, V - 1, decode=True, ctc_bounds=bounds)
    torch.cuda.synchronize()
    assert _lib.lib().asrk_launch_count() - n0 == 3          # transform, statistics, fused CTC + z-score
    assert torch.equal(f, f_ref)
    assert torch.equal(r.loss, r_ref.loss) and torch.equal(r.grad, r_ref.grad) and torch.equal(r.row_status, r_ref.row_status)
    if kind == "logits":
        assert torch.equal(r.token_len, r_ref.token_len)
        assert ctc.tokens_to_lists(r.tokens, r.token_len) == ctc.tokens_to_lists(r_ref.tokens, r_ref.token_len)
        # a batch whose lattices do not fit the fused kernel: same call, two-kernel tail
        f2, r2 = step(samples, so, sc, fo, B, nf, logits, dlab, dll, dil, V - 1, decode=True, ctc_bounds=(5000, 600))
        torch.cuda.synchronize()
        assert torch.equal(f2, f_ref) and torch.equal(r2.loss, r_ref.loss) and torch.equal(r2.grad, r_ref.grad)


@pytest.mark.parametrize("lanes,feature_ctas,merged,generic", [(2, 104, False, False), (3, 40, False, False),
                                                               (2, 104, True, False), (2, 104, False, True)])
def test_steps_in_flight_equal_serial_steps(lanes, feature_ctas, merged, generic):
    """pipeline.StepsInFlight: the captured steps of different batches replayed round-robin on several lane streams
    (the next batch's transform next to the previous batch's HBM-bound tail) return bit for bit what the serial step
    returns -- features, loss, gradient, tokens, [sum loss, n] -- over several rounds with poisoned output buffers;
    with the fused CTC kernel, the merged tail and the generic long-lattice kernels."""
    from asr_dfcnn_transformer_b200 import ctc, features, pipeline
    dev = torch.device("cuda", 0)
    n_batches = 2 * lanes
    serial = pipeline.HotPathStep(dev)
    flight = pipeline.StepsInFlight(dev, lanes=lanes, feature_ctas=feature_ctas, merged_tail=merged)
    want, slots, outs = [], [], []
    for i in range(n_batches):
        pcm, x, labels, ll, il = _batch(1300 + i, B=40 + 9 * i)
        B, V = x.shape[1:]
        pk = features.pack_host(pcm, pin=False)
        so = torch.as_tensor(np.asarray(pk.sample_offsets, dtype=np.int64)).to(dev)
        sc = torch.as_tensor(np.asarray(pk.sample_counts, dtype=np.int64)).to(dev)
        fo = torch.as_tensor(np.asarray(pk.frame_offsets, dtype=np.int64)).to(dev)
        nf = int(pk.frame_offsets[-1])
        a = (pk.samples.to(dev), so, sc, fo, B, nf, torch.as_tensor(x).to(dev), torch.as_tensor(labels.astype(np.int32)).to(dev),
             torch.as_tensor(ll).to(dev), torch.as_tensor(il).to(dev), V - 1)
        # generic: bounds beyond the fused kernel's lattices -> rows / lattice / grad / collapse kernels (the C3 path)
        bounds = (5000, 600) if generic else (int(il.max()), int(ll.max()))
        f, r = serial(*a, decode=True, ctc_bounds=bounds)
        torch.cuda.synchronize()
        want.append((f.clone(), r.loss.clone(), r.grad.clone(), ctc.tokens_to_lists(r.tokens, r.token_len),
                     ctc.loss_sum(r.loss, r.row_status).clone()))
        feat = torch.empty((nf, 200), dtype=torch.float32, device=dev)
        grad = torch.empty_like(a[6])
        outs.append((feat, grad))
        slots.append(flight.add(*a, feat_out=feat, grad_out=grad, decode=True, ctc_bounds=bounds))
    for rnd in range(3):
        for feat, grad in outs:
            feat.fill_(float("nan"))
            grad.fill_(float("nan"))
        torch.cuda.synchronize()
        for s in slots:
            flight.launch(s)
        flight.join()
        torch.cuda.synchronize()
        for s, (f, loss, grad, toks, lsum) in zip(slots, want):
            assert torch.equal(s.features, f), rnd
            assert torch.equal(s.result.loss, loss) and torch.equal(s.result.grad, grad), rnd
            assert ctc.tokens_to_lists(s.result.tokens, s.result.token_len) == toks
            assert torch.equal(s.loss_sum, lsum)
