"""GPU parity: CTC loss / gradient and greedy decode vs the float64 oracle."""
import numpy as np
import pytest

from oracle import ctc_ref, synth
from tests.util import CTC_ATOL, CTC_RTOL, assert_ctc_grad_close

pytestmark = pytest.mark.gpu


def _run(x, labels, ll, il, blank=None, label_mode="by_length", layout="tbv", decode=False):
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    xt = torch.as_tensor(x).cuda()
    if layout == "btv":
        xt = xt.permute(1, 0, 2).contiguous()
    r = ctc.ctc_loss_grad(xt, labels, ll, il, blank, label_mode, layout, decode=decode)
    g = r.grad
    if layout == "btv":
        g = g.permute(1, 0, 2)
    return r, r.loss.cpu().numpy(), g.cpu().numpy(), r.row_status.cpu().numpy()


def _check(x, labels, ll, il, blank, **kw):
    r, loss, grad, status = _run(x, labels, ll, il, blank, **kw)
    mode = kw.get("label_mode", "by_length")
    rl, rg, ok = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, blank, label_mode=mode)
    assert np.all(status[ok] == 0)
    np.testing.assert_allclose(loss[ok], rl[ok], rtol=CTC_RTOL, atol=CTC_ATOL)
    assert_ctc_grad_close(grad, rg, x, il)
    return r


def test_small_cases_tbv_and_btv():
    rng = np.random.default_rng(0)
    x, labels, ll, il = synth.ctc_batch(rng, [20, 13, 7, 30, 1, 2], 52, 0, 9)
    _check(x, labels, ll, il, 51)
    _check(x, labels, ll, il, 51, layout="btv")


def test_known_answers():
    # T=1, L=0: loss = -log softmax(blank)
    x = np.array([[[0.0, 1.0, 2.0, 3.0]]], dtype=np.float32)            # [T=1,B=1,V=4]
    r, loss, grad, st = _run(x, np.zeros((1, 1), np.int32), [0], [1], 3)
    lp = x[0, 0] - np.log(np.exp(x[0, 0]).sum())
    assert abs(loss[0] + lp[3]) < 1e-5
    y = np.exp(lp); y[3] -= 1.0
    np.testing.assert_allclose(grad[0, 0], y, atol=1e-5)
    # T=2, L=1 (label 0): paths "0b", "b0", "00"
    x = np.zeros((2, 1, 3), dtype=np.float32)
    r, loss, grad, st = _run(x, np.array([[0]], np.int32), [1], [2], 2)
    assert abs(loss[0] + np.log(3.0 / 9.0)) < 1e-5
    # repeated label needs T = L + repeats exactly: labels (1,1), T=3 -> single path "1 b 1"
    x = np.zeros((3, 1, 3), dtype=np.float32)
    r, loss, grad, st = _run(x, np.array([[1, 1]], np.int32), [2], [3], 2)
    assert abs(loss[0] + 3 * np.log(1.0 / 3.0)) < 1e-5 and st[0] == 0
    # not enough time: labels (1,1), T=2
    r, loss, grad, st = _run(x[:2], np.array([[1, 1]], np.int32), [2], [2], 2)
    assert st[0] == 2 and np.isinf(loss[0])
    np.testing.assert_allclose(grad[:, 0], np.full((2, 3), 1.0 / 3.0), atol=1e-6)   # dy = y


def test_label_zero_and_drop_zeros_mode():
    rng = np.random.default_rng(1)
    x, labels, ll, il = synth.ctc_batch(rng, [25, 25, 18], 40, 4, 8, lmax=12)
    labels[0, 1] = 0      # a genuine label id 0
    _check(x, labels, ll, il, 39)
    _check(x, labels, ll, il, 39, label_mode="drop_zeros")


def test_input_len_shorter_than_T_zero_grad_rows():
    rng = np.random.default_rng(2)
    x, labels, ll, il = synth.ctc_batch(rng, [9, 30, 17], 64, 2, 6)
    r, loss, grad, st = _run(x, labels, ll, il, 63)
    assert np.all(grad[9:, 0] == 0) and np.all(grad[17:, 2] == 0)
    # sum_v grad = 0 on valid frames (softmax minus a distribution)
    assert np.abs(grad[:9, 0].sum(-1)).max() < 1e-4


def test_c1_shape_dict_vocab():
    c = synth.config_c1()
    _check(c["logits"], c["labels"], c["label_len"], c["input_len"], c["V"] - 1)


def test_c2_shape_subset():
    rng = np.random.default_rng(2000)
    il = np.array([synth.t_ctc(synth.n_frames(int(n))) for n in synth.ragged_lengths(rng, 24, 3.0, 7.0)], np.int32)
    x, labels, ll, il = synth.ctc_batch(rng, il, synth.VOCAB_DICT_TXT, 8, 24, lmax=64)
    _check(x, labels, ll, il, synth.VOCAB_DICT_TXT - 1)


def test_long_lattice_c3_shape_one_row():
    rng = np.random.default_rng(3000)
    x, labels, ll, il = synth.ctc_batch(rng, [1998, 1500], synth.VOCAB_DICT_TXT, 280, 320)
    _check(x, labels, ll, il, synth.VOCAB_DICT_TXT - 1)


def test_mixdict_vocab_and_unaligned_vocab():
    rng = np.random.default_rng(4)
    for V in (synth.VOCAB_MIXDICT, 1423, 37):
        x, labels, ll, il = synth.ctc_batch(rng, [40, 33, 12], V, 3, 10)
        _check(x, labels, ll, il, V - 1)


def _keras_case(rng, T_list, V, lo, hi, lmax):
    import torch
    x, labels, ll, il = synth.ctc_batch(rng, T_list, V, lo, hi, lmax=lmax)
    p = torch.softmax(torch.as_tensor(x).permute(1, 0, 2), -1).contiguous()      # [B,T,V]
    return p, labels, ll, il


def _assert_prob_grad_close(got, ref, p, il):
    """Gradient w.r.t. y_pred: dL/dp = (y - occupancy) / (p + eps).  "Within 1e-3 relative" on the two
    probability terms, i.e. |d| (p + eps) <= rtol (|y - occ| + occ) + atol (the same rule as
    assert_ctc_grad_close, carried through the division)."""
    pe = p.astype(np.float64) + 1e-7
    y = pe / pe.sum(-1, keepdims=True)
    B, T, _ = p.shape
    valid = (np.arange(T)[None, :] < np.asarray(il)[:, None])[:, :, None]
    dref = ref * pe                               # y - occ
    occ = np.abs(y * valid - dref)
    tol = CTC_RTOL * (np.abs(dref) + occ) + CTC_ATOL
    bad = np.abs(got - ref) * pe > tol
    assert not bad.any(), (int(bad.sum()), float((np.abs(got - ref) * pe)[bad].max()))
    assert not got[~np.broadcast_to(valid, got.shape)].any()


def test_keras_ctc_batch_cost_and_autograd():
    """K.ctc_batch_cost in ONE kernel: log(y_pred + 1e-7) formed on load, gradient written w.r.t. y_pred,
    upstream scale inside (cnn_ctc.py:149-152).  Small lattices (fused kernel) and a long one (generic
    kernels); `.sum().backward()` and the batch mean with the scale stated up front."""
    import torch
    from asr_dfcnn_transformer_b200 import _lib, ctc
    rng = np.random.default_rng(5)
    for T_list, V, lo, hi, lmax in (([30, 22, 16, 30], 60, 3, 9, 16), ([63, 40, 88, 12], 1424, 8, 24, 64),
                                    ([300, 150], 64, 40, 60, 64)):
        p, labels, ll, il = _keras_case(rng, T_list, V, lo, hi, lmax)
        B = len(il)
        args = (torch.as_tensor(labels.astype(np.float32)), None, torch.as_tensor(il.reshape(-1, 1).astype(np.int64)),
                torch.as_tensor(ll.reshape(-1, 1).astype(np.int64)))
        rl, rgp = ctc_ref.keras_ctc_batch_cost(labels, p.numpy(), il, ll)
        # (a) .sum(): exactly one kernel of the library for forward + backward
        pt = p.cuda().requires_grad_(True)
        torch.cuda.synchronize()
        n0 = _lib.lib().asrk_launch_count()
        cost = ctc.ctc_batch_cost(args[0], pt, args[2], args[3])
        assert tuple(cost.shape) == (B, 1)
        cost.sum().backward()
        n1 = _lib.lib().asrk_launch_count()
        if max(T_list) <= 88:
            assert n1 - n0 == 1, n1 - n0            # the fused kernel, nothing else
        np.testing.assert_allclose(cost.detach().cpu().numpy(), rl, rtol=CTC_RTOL, atol=CTC_ATOL)
        _assert_prob_grad_close(pt.grad.cpu().numpy().astype(np.float64), rgp, p.numpy(), il)
        # (b) Keras' batch mean with the upstream gradient stated: still no extra pass, gradient scaled by 1/B
        pt2 = p.cuda().requires_grad_(True)
        ctc.ctc_batch_cost(args[0], pt2, args[2], args[3], upstream=1.0 / B).mean().backward()
        _assert_prob_grad_close(pt2.grad.cpu().numpy().astype(np.float64) * B, rgp, p.numpy(), il)
        # (c) a non-uniform upstream gradient falls back to one element-wise pass and is still right
        pt3 = p.cuda().requires_grad_(True)
        w = torch.arange(1, B + 1, dtype=torch.float32, device="cuda").view(B, 1)
        (ctc.ctc_batch_cost(args[0], pt3, args[2], args[3]) * w).sum().backward()
        _assert_prob_grad_close(pt3.grad.cpu().numpy().astype(np.float64) / w.cpu().numpy().reshape(B, 1, 1), rgp,
                                p.numpy(), il)


def test_ctc_loss_v2_blank_default_like_tf():
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    rng = np.random.default_rng(15)
    x, labels, ll, il = synth.ctc_batch(rng, [20, 14], 12, 2, 5)
    labels = labels + 1                                   # blank 0: labels in 1..V-1
    labels[labels >= 12] = 11
    xs = torch.as_tensor(x).cuda()
    loss = ctc.ctc_loss_v2(torch.as_tensor(labels), xs, torch.as_tensor(ll), torch.as_tensor(il))   # blank_index=None -> 0
    rl, _, _ = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, 0)
    np.testing.assert_allclose(loss.cpu().numpy(), rl, rtol=CTC_RTOL, atol=CTC_ATOL)
    with pytest.raises(ValueError):
        ctc.ctc_loss_v2(ctc.dense_to_sparse(torch.as_tensor(labels)), xs, None, torch.as_tensor(il))
    # T == 0: every row rejected, nothing undefined
    r = ctc.ctc_loss_grad(torch.zeros((0, 2, 12), device="cuda"), labels, ll, il, 11)
    assert (r.row_status.cpu().numpy() == 3).all() and np.isnan(r.loss.cpu().numpy()).all()


def test_ctc_loss_v2_raises_like_tf():
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    x = torch.zeros((2, 1, 3), device="cuda")
    with pytest.raises(ctc.InvalidArgumentError):
        ctc.ctc_loss_v2(torch.tensor([[1, 1]], dtype=torch.int32), x, torch.tensor([2]), torch.tensor([2]),
                        blank_index=2)


def test_greedy_decode_bit_exact():
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    rng = np.random.default_rng(6)
    for V, scale in ((synth.VOCAB_DICT_TXT, 3.0), (50, 0.5), (7, 0.2)):
        il = np.array([63, 1, 40, 88, 33, 64, 32, 31], np.int32)
        x = synth.logits_tbv(rng, 88, len(il), V, scale)
        # force ties, repeats and blanks
        x[5:9, 0, :] = 0.0                 # all equal -> index 0 four times
        x[10:13, 2, V - 1] = 50.0          # blanks
        x[13:15, 2, 3] = 50.0              # repeat "3 3"
        x[15, 2, V - 1] = 50.0
        x[16, 2, 3] = 50.0                 # "3" again after a blank -> emitted twice
        x = np.round(x * 4) / 4            # many exact ties
        tokens, tlen, nsl = ctc.greedy_decode(torch.as_tensor(x).cuda(), il)
        got = ctc.tokens_to_lists(tokens, tlen)
        ref, rnsl = ctc_ref.greedy_decode(x, il)
        assert got == ref
        np.testing.assert_allclose(nsl.cpu().numpy(), rnsl, rtol=1e-5, atol=1e-4)
        got2, _ = ctc_ref.greedy_decode(x, il, merge_repeated=False)
        t2, l2, _ = ctc.greedy_decode(torch.as_tensor(x).cuda(), il, merge_repeated=False)
        assert ctc.tokens_to_lists(t2, l2) == got2


def test_fused_decode_matches_standalone():
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    rng = np.random.default_rng(8)
    x, labels, ll, il = synth.ctc_batch(rng, [50, 41, 63, 9], 200, 3, 12)
    r, *_ = _run(x, labels, ll, il, 199, decode=True)
    ref, _ = ctc_ref.greedy_decode(x, il)
    assert ctc.tokens_to_lists(r.tokens, r.token_len) == ref


def test_tf_decoder_surface_and_decode_ctc():
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    rng = np.random.default_rng(9)
    x = synth.logits_tbv(rng, 30, 2, 20, 1.0)
    dec, nsl = ctc.ctc_greedy_decoder(torch.as_tensor(x).cuda(), np.array([30, 12], np.int32))
    ref, _ = ctc_ref.greedy_decode(x, [30, 12])
    dense = ctc.sparse_tensor_to_dense(dec[0], default_value=0)
    assert np.array_equal(dense, ctc_ref.densify(ref, 0))
    assert nsl.shape == (2, 1)
    p = np.exp(x[:, :1]) / np.exp(x[:, :1]).sum(-1, keepdims=True)
    ids = ctc.decode_ctc(np.transpose(p, (1, 0, 2)).astype(np.float32), 30)
    ref1, _ = ctc_ref.greedy_decode(np.log(p.astype(np.float32) + np.float32(1e-7)), [30])
    assert ids.tolist() == ref1[0]


def test_full_size_c2_batch_properties():
    """BASELINE.json configs[1] at full size: properties that do not need the oracle on all
    of it -- every valid gradient row sums to 0 (softmax minus a distribution), rows past
    input_len are exactly 0, the loss is finite and positive, and the batch reduction kernel
    agrees with a float64 sum; a slice of utterances is checked against the oracle."""
    import torch
    import bench
    from asr_dfcnn_transformer_b200 import ctc
    hb = bench.make_batch(2001)
    x, labels, ll, il = hb["logits"], hb["labels"], hb["label_len"], hb["input_len"]
    r = ctc.ctc_loss_grad(torch.as_tensor(x).cuda(), labels, ll, il, x.shape[2] - 1, decode=True)
    loss = r.loss.cpu().numpy()
    grad = r.grad.cpu().numpy()
    assert int(r.row_status.max()) == 0 and np.all(np.isfinite(loss)) and np.all(loss > 0)
    T = x.shape[0]
    valid = np.arange(T)[:, None] < il[None, :]
    assert np.abs(grad.sum(-1))[valid].max() < 1e-4     # 10x inside the 1e-3 tolerance
    assert not grad[~valid].any()
    red = ctc.loss_sum(r.loss, r.row_status).cpu().numpy()
    assert red[1] == 256 and abs(red[0] - loss.astype(np.float64).sum()) < 1e-6 * red[0]
    # run-to-run determinism: the per-warp row buffers are refilled by TMA bulk copies while other
    # warps compute; a buffer handed back before its loads had landed showed up as a few rows in
    # 10^4 differing between runs
    xs = torch.as_tensor(x).cuda()
    for _ in range(8):
        r2 = ctc.ctc_loss_grad(xs, labels, ll, il, x.shape[2] - 1, decode=True)
        assert torch.equal(r2.grad, r.grad) and torch.equal(r2.loss, r.loss)
    sl = slice(40, 56)
    rl, rg, ok = ctc_ref.ctc_loss_grad_batch(np.ascontiguousarray(x[:, sl]), labels[sl], ll[sl], il[sl], x.shape[2] - 1)
    np.testing.assert_allclose(loss[sl], rl, rtol=CTC_RTOL, atol=CTC_ATOL)
    assert_ctc_grad_close(grad[:, sl], rg, x[:, sl], il[sl])
    ref_tok, _ = ctc_ref.greedy_decode(x, il)
    assert ctc.tokens_to_lists(r.tokens, r.token_len) == ref_tok


def test_edge_shapes():
    """T = 1, empty label, single class besides blank, batch of one, label length == T-repeats."""
    rng = np.random.default_rng(9)
    x, labels, ll, il = synth.ctc_batch(rng, [1, 1, 5, 4], 6, 0, 1)
    ll[0] = 0                                   # empty label with one frame: p = y(blank)
    _check(x, labels, ll, il, 5)
    x = rng.standard_normal((7, 1, 2)).astype(np.float32)
    _check(x, np.array([[0, 0, 0]]), np.array([3]), np.array([5]), 1)      # "0 _ 0 _ 0": the only path


def test_full_size_c3_long_utterance_batch_properties():
    """BASELINE.json configs[2] at full size: 64 utterances, T = 1998 frames, ~300 labels (the
    generic row / lattice / gradient kernels).  Properties on all of it, the oracle on one row."""
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    rng = np.random.default_rng(3001)
    V = synth.VOCAB_DICT_TXT
    il = np.full(64, 1998, dtype=np.int32)
    il[5], il[17] = 1500, 1001
    x, labels, ll, il = synth.ctc_batch(rng, il, V, 280, 320)
    r = ctc.ctc_loss_grad(torch.as_tensor(x).cuda(), labels, ll, il, V - 1, decode=True)
    loss = r.loss.cpu().numpy()
    assert int(r.row_status.max()) == 0 and np.all(np.isfinite(loss)) and np.all(loss > 0)
    T = x.shape[0]
    valid = np.arange(T)[:, None] < il[None, :]
    s = torch.abs(r.grad.sum(-1)).cpu().numpy()
    assert s[valid].max() < 1e-4
    assert not s[~valid].any()
    b = 17
    rl, rg, ok = ctc_ref.ctc_loss_grad_batch(np.ascontiguousarray(x[:, b:b + 1]), labels[b:b + 1], ll[b:b + 1],
                                             il[b:b + 1], V - 1)
    np.testing.assert_allclose(loss[b], rl[0], rtol=CTC_RTOL, atol=CTC_ATOL)
    assert_ctc_grad_close(r.grad[:, b:b + 1].cpu().numpy(), rg, x[:, b:b + 1], il[b:b + 1])
    ref_tok, _ = ctc_ref.greedy_decode(x[:, :8], il[:8])
    assert ctc.tokens_to_lists(r.tokens, r.token_len)[:8] == ref_tok


def test_stage_logits_copies_only_valid_rows():
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    rng = np.random.default_rng(12)
    for layout, shape in (("tbv", (9, 5, 64)), ("btv", (5, 9, 64))):
        h = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).pin_memory()
        il = torch.tensor([9, 1, 4, 7, 2], dtype=torch.int32, device="cuda")
        out = torch.full(shape, -7.0, dtype=torch.float32, device="cuda")
        ctc.stage_logits(h, il, out=out, layout=layout)
        got, ref = out.cpu().numpy(), h.numpy()
        if layout == "btv":
            got, ref = got.transpose(1, 0, 2), ref.transpose(1, 0, 2)
        for b, n in enumerate([9, 1, 4, 7, 2]):
            assert np.array_equal(got[:n, b], ref[:n, b])
            assert np.all(got[n:, b] == -7.0)              # padding rows never touched
    with pytest.raises(ValueError):
        ctc.stage_logits(torch.zeros(2, 2, 4), il[:2])


def test_bounded_batch_is_one_launch_and_matches_the_unbounded_call():
    """Lengths given as DEVICE tensors leave the host blind: prep + fused + the generic kernels are
    launched.  With ``bounds`` (or host length vectors) the fused kernel prepares its own utterance
    and nothing else runs.  Both must give the same bits; a broken promise is reported per row."""
    import torch
    from asr_dfcnn_transformer_b200 import _lib, ctc
    rng = np.random.default_rng(77)
    il = rng.integers(20, 60, 12).astype(np.int32)
    x, labels, ll, il = synth.ctc_batch(rng, il, 1424, 4, 20, lmax=64)
    xs = torch.as_tensor(x).cuda()
    d = lambda a: torch.as_tensor(a).cuda()
    a = ctc.ctc_loss_grad(xs, d(labels), d(ll), d(il), 1423, decode=True)               # unbounded
    b = ctc.ctc_loss_grad(xs, d(labels), d(ll), d(il), 1423, decode=True, bounds=(int(il.max()), int(ll.max())))
    c = ctc.ctc_loss_grad(xs, labels, ll, il, 1423, decode=True)                        # host vectors
    for r in (b, c):
        assert torch.equal(r.loss, a.loss) and torch.equal(r.grad, a.grad)
        assert torch.equal(r.row_status, a.row_status) and torch.equal(r.token_len, a.token_len)
        for k in range(len(il)):
            n = int(a.token_len[k])
            assert torch.equal(r.tokens[k, :n], a.tokens[k, :n])
    ref = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, 1423)
    assert np.allclose(a.loss.cpu().numpy(), ref[0], rtol=CTC_RTOL, atol=CTC_ATOL)
    # a promise that does not hold: rows with more than 31 labels cannot take the fused kernel
    il2 = np.full(4, 120, np.int32)
    x2, labels2, ll2, il2 = synth.ctc_batch(rng, il2, 1424, 10, 50, lmax=64)
    ll2[1] = 40
    labels2[1, :40] = 1 + np.arange(40)
    r = ctc.ctc_loss_grad(torch.as_tensor(x2).cuda(), d(labels2), d(ll2), d(il2), 1423, bounds=(60, 20))
    st = r.row_status.cpu().numpy()
    assert st[1] == _lib.ROW_NOT_SMALL and np.isnan(r.loss.cpu().numpy()[1])
    assert not r.grad[:, 1].any()
    with pytest.raises(ValueError):
        ctc._raise_on_status(r.row_status)
    # without the promise the same batch goes through the generic kernels
    r2 = ctc.ctc_loss_grad(torch.as_tensor(x2).cuda(), labels2, ll2, il2, 1423)
    ref2 = ctc_ref.ctc_loss_grad_batch(x2, labels2, ll2, il2, 1423)
    assert int(r2.row_status.max()) == 0
    assert np.allclose(r2.loss.cpu().numpy(), ref2[0], rtol=CTC_RTOL, atol=CTC_ATOL)
