"""Property tests (hypothesis): size-independent properties of the three parts, on the oracle (CPU) and on the CUDA
kernels (GPU).  SURVEY.md section 4: the gradient of a valid frame sums to 0 and frames past input_len are exactly 0;
the loss is >= 0; the greedy decode is idempotent on its own output; z-scored feature columns have mean 0, std 1."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import ctc_ref, fbank_ref, synth

ctc_case = st.tuples(st.integers(1, 6), st.integers(1, 40), st.integers(3, 50), st.integers(0, 2 ** 31 - 1))


def _case(B, T, V, seed):
    rng = np.random.default_rng(seed)
    il = rng.integers(1, T + 1, B).astype(np.int32)
    il[rng.integers(0, B)] = T
    x, labels, ll, il = synth.ctc_batch(rng, il, V, 0, max(0, min(T - 1, 12)))
    return x, labels, ll, il


def _check_ctc_properties(x, il, loss, grad, ok):
    T = x.shape[0]
    valid = (np.arange(T)[:, None] < il[None, :]) & ok[None, :]
    assert np.all(loss[ok] >= -1e-6)
    s = np.abs(grad.sum(-1))
    assert s[valid].max(initial=0.0) < 1e-4
    assert not grad[~(np.arange(T)[:, None] < il[None, :])].any()
    assert np.all(grad[valid] <= 1.0 + 1e-5) and np.all(grad[valid] >= -1.0 - 1e-5)


@settings(max_examples=40, deadline=None)
@given(ctc_case)
def test_oracle_ctc_properties(c):
    B, T, V, seed = c
    x, labels, ll, il = _case(B, T, V, seed)
    loss, grad, ok = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, V - 1)
    _check_ctc_properties(x, il, loss, grad, ok)


@settings(max_examples=40, deadline=None)
@given(ctc_case)
def test_oracle_decode_is_idempotent(c):
    B, T, V, seed = c
    x, labels, ll, il = _case(B, T, V, seed)
    tok, _ = ctc_ref.greedy_decode(x, il)
    # one-hot logits of the decoded sequence (with a blank between equal neighbours) decode to themselves
    for b, seq in enumerate(tok):
        path = []
        for k, c_ in enumerate(seq):
            if k and seq[k - 1] == c_:
                path.append(V - 1)
            path.append(c_)
        if not path:
            continue
        y = np.full((len(path), 1, V), -5.0, np.float32)
        y[np.arange(len(path)), 0, path] = 5.0
        again, _ = ctc_ref.greedy_decode(y, [len(path)])
        assert again[0] == seq
        assert V - 1 not in seq and all(0 <= c_ < V - 1 for c_ in seq)


@settings(max_examples=15, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(800, 20000))
def test_oracle_feature_columns_are_standardised(seed, n):
    rng = np.random.default_rng(seed)
    sig = synth.g2_voiced(rng, n) if seed % 2 else synth.g1_white(rng, n)
    fb = fbank_ref.compute_fbank(sig)
    assert fb.shape == (fbank_ref.n_frames_fbank(n), 200)
    if fb.shape[0] >= 2:
        assert np.abs(fb.mean(axis=0)).max() < 1e-9
        sd = fb.std(axis=0)
        assert np.all((np.abs(sd - 1.0) < 1e-9) | (sd < 1e-12))


@pytest.mark.gpu
@settings(max_examples=25, deadline=None)
@given(ctc_case)
def test_gpu_ctc_properties(c):
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    B, T, V, seed = c
    x, labels, ll, il = _case(B, T, V, seed)
    r = ctc.ctc_loss_grad(torch.as_tensor(x).cuda(), labels, ll, il, V - 1, decode=True)
    st_ = r.row_status.cpu().numpy()
    ok = st_ == 0
    _check_ctc_properties(x, il, r.loss.cpu().numpy(), r.grad.cpu().numpy(), ok)
    ref_tok, _ = ctc_ref.greedy_decode(x, il)
    assert ctc.tokens_to_lists(r.tokens, r.token_len) == ref_tok


@pytest.mark.gpu
@settings(max_examples=10, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.lists(st.integers(560, 30000), min_size=1, max_size=5))
def test_gpu_feature_columns_are_standardised(seed, lens):
    from asr_dfcnn_transformer_b200 import features
    rng = np.random.default_rng(seed)
    sigs = [synth.g2_voiced(rng, n) if (seed + i) % 2 else synth.g1_white(rng, n) for i, n in enumerate(lens)]
    fb = features.compute_features(sigs, mode="fbank")
    out = fb.features.cpu().numpy().astype(np.float64)
    for i, n in enumerate(lens):
        y = out[fb.frame_offsets[i]:fb.frame_offsets[i + 1]]
        assert y.shape[0] == fbank_ref.n_frames_fbank(n)
        if y.shape[0] >= 2:
            assert np.abs(y.mean(axis=0)).max() < 2e-4
            sd = y.std(axis=0)
            assert np.all((np.abs(sd - 1.0) < 2e-3) | (sd == 0.0))
