"""GPU parity: LFR stacking and the label-error (edit distance) kernel vs the oracle."""
import numpy as np
import pytest

from oracle import utils_ref

pytestmark = pytest.mark.gpu


def test_lfr_matches_reference_loop():
    import torch
    from asr_dfcnn_transformer_b200 import utils
    rng = np.random.default_rng(0)
    for T, m, n in [(7, 4, 3), (498, 4, 3), (1, 4, 3), (3, 4, 3), (12, 1, 1), (10, 1, 3), (9, 5, 1)]:
        x = rng.standard_normal((T, 200)).astype(np.float32)
        got = utils.build_LFR_features(x, m, n)
        assert np.array_equal(got, utils_ref.build_LFR_features(x, m, n)), (T, m, n)     # a pure gather: bit-exact
    # ragged batch in one launch
    Ts = [5, 498, 1, 33]
    xs = [rng.standard_normal((t, 200)).astype(np.float32) for t in Ts]
    fo = np.concatenate([[0], np.cumsum(Ts)])
    out, oo = utils.lfr_batch(torch.from_numpy(np.concatenate(xs)).cuda(), fo, 4, 3)
    out = out.cpu().numpy()
    for i, x in enumerate(xs):
        assert np.array_equal(out[oo[i]:oo[i + 1]], utils_ref.build_LFR_features(x, 4, 3))


def test_edit_distance_matches_oracle():
    import torch
    from asr_dfcnn_transformer_b200 import utils
    rng = np.random.default_rng(1)
    B, Hmax, Lmax = 300, 90, 64
    hyp = rng.integers(0, 6, (B, Hmax)).astype(np.int32)
    truth = rng.integers(0, 6, (B, Lmax)).astype(np.int32)
    hl = rng.integers(0, Hmax + 1, B).astype(np.int32)
    tl = rng.integers(0, Lmax + 1, B).astype(np.int32)
    hl[:4] = [0, 0, 5, Hmax]
    tl[:4] = [0, 7, 0, Lmax]
    for normalize in (True, False):
        got = utils.edit_distance(torch.from_numpy(hyp).cuda(), hl, truth, tl, normalize=normalize).cpu().numpy()
        ref = utils_ref.edit_distance([hyp[b, :hl[b]] for b in range(B)], [truth[b, :tl[b]] for b in range(B)], normalize)
        assert np.array_equal(np.isinf(got), np.isinf(ref))
        fin = ~np.isinf(ref)
        np.testing.assert_allclose(got[fin], ref[fin], rtol=1e-6)


def test_label_error_of_greedy_decode():
    """The chain the reference runs: greedy decode -> edit distance against the labels -> mean
    (acoustic_model2.py:69-73), entirely on the device."""
    import torch
    from asr_dfcnn_transformer_b200 import ctc, utils
    from oracle import ctc_ref, synth
    rng = np.random.default_rng(2)
    x, labels, ll, il = synth.ctc_batch(rng, [40, 33, 25, 12, 60], 30, 3, 9, lmax=16)
    r = ctc.ctc_loss_grad(torch.as_tensor(x).cuda(), labels, ll, il, decode=True)
    ler = float(utils.label_error_rate(r.tokens, r.token_len, labels, ll))
    toks, _ = ctc_ref.greedy_decode(x, il)
    ref = utils_ref.edit_distance(toks, [labels[b, :ll[b]] for b in range(len(ll))]).mean()
    assert abs(ler - ref) < 1e-6
    assert utils.GetEditDistance("abcd", "abed") == 1 and utils.GetEditDistance("", "xy") == 2
