"""CPU: the CTC / decode oracles against each other, torch float64 and known answers."""
import numpy as np
import pytest

from oracle import build_c, ctc_ref, synth


def test_scalar_vs_vectorised_vs_torch_float64():
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(0)
    x, labels, ll, il = synth.ctc_batch(rng, [20, 13, 7, 30, 1, 2, 9], 50, 0, 9)
    l1, g1, _ = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, 49, impl="loop")
    l2, g2, _ = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, 49, impl="vec")
    assert np.abs(l1 - l2).max() < 1e-10 and np.abs(g1 - g2).max() < 1e-10
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    lt = torch.nn.functional.ctc_loss(torch.log_softmax(xt, -1), torch.tensor(labels, dtype=torch.long),
                                      torch.tensor(il, dtype=torch.long), torch.tensor(ll, dtype=torch.long),
                                      blank=49, reduction="none", zero_infinity=False)
    lt.sum().backward()
    assert np.abs(lt.detach().numpy() - l1).max() < 1e-10
    assert np.abs(xt.grad.numpy() - g1).max() < 1e-10


def test_known_answers():
    # T=2, L=1, uniform over V=3: p = P("0b") + P("b0") + P("00") = 3/9
    x = np.zeros((2, 1, 3), dtype=np.float32)
    l, g, ok = ctc_ref.ctc_loss_grad_batch(x, np.array([[0]]), [1], [2], 2)
    assert ok[0] and abs(l[0] + np.log(3 / 9)) < 1e-12
    assert np.abs(g.sum(-1)).max() < 1e-12            # softmax minus a distribution
    # repeated label needs a blank in between: (1,1) with T=3 -> exactly "1 b 1"
    x = np.zeros((3, 1, 3), dtype=np.float32)
    l, g, ok = ctc_ref.ctc_loss_grad_batch(x, np.array([[1, 1]]), [2], [3], 2)
    assert ok[0] and abs(l[0] + 3 * np.log(1 / 3)) < 1e-12
    # ... and T=2 is infeasible: +inf loss, dy = y
    l, g, ok = ctc_ref.ctc_loss_grad_batch(x[:2], np.array([[1, 1]]), [2], [2], 2)
    assert not ok[0] and np.isinf(l[0]) and np.allclose(g, 1 / 3)
    # frames beyond input_len get zero gradient
    rng = np.random.default_rng(1)
    x, labels, ll, il = synth.ctc_batch(rng, [9, 20], 12, 2, 4)
    l, g, ok = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, 11)
    assert np.all(g[9:, 0] == 0)


def test_label_modes():
    rng = np.random.default_rng(2)
    x, labels, ll, il = synth.ctc_batch(rng, [25, 25], 30, 5, 8, lmax=10)
    labels[0, 2] = 0
    la, _, _ = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, 29, label_mode="by_length")
    lb, _, _ = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, 29, label_mode="drop_zeros")
    assert la[0] != lb[0]          # acoustic_model2.py:71 drops the genuine label 0


def test_keras_glue_shift():
    # TF re-normalises log(p + eps): the loss moves by T*log(1 + V*eps)
    rng = np.random.default_rng(3)
    #  when every p >> eps
    x, labels, ll, il = synth.ctc_batch(rng, [40], 1424, 5, 9, scale=0.5)
    p = np.exp(ctc_ref.log_softmax(np.transpose(x, (1, 0, 2))))
    lk, _ = ctc_ref.keras_ctc_batch_cost(labels, p, il, ll)
    l0, _, _ = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, 1423)
    assert lk.shape == (1, 1)
    # the two effects of eps (+eps/p per frame, -log(1 + V eps) per frame) are each
    # bounded by T*log(1 + V*eps) here
    shift = 40 * np.log1p(1424 * 1e-7)
    assert 0 < abs(lk[0, 0] - l0[0]) < shift


def test_c_restatement_matches_numpy_oracle():
    rng = np.random.default_rng(4)
    x, labels, ll, il = synth.ctc_batch(rng, [40, 33, 12, 25, 1], 200, 0, 10)
    rl, rg, ok = ctc_ref.ctc_loss_grad_batch(x, labels, ll, il, 199)
    l64, g64, s64 = build_c.ctc_loss_grad(x, labels, ll, il, 199, real="f64")
    assert np.abs(l64 - rl).max() < 1e-4 and np.abs(g64 - rg).max() < 1e-6 and not s64.any()
    l32, g32, s32 = build_c.ctc_loss_grad(x, labels, ll, il, 199, real="f32")
    assert np.allclose(l32, rl, rtol=1e-4) and np.abs(g32 - rg).max() < 1e-3
    seqs, nsl = build_c.greedy_decode(x, il, 199)
    ref, rnsl = ctc_ref.greedy_decode(x, il)
    assert seqs == ref and np.allclose(nsl, rnsl, rtol=1e-5)


def test_greedy_decode_rules():
    V = 5
    x = np.full((6, 1, V), -5.0, dtype=np.float32)
    for t, c in enumerate([1, 1, 4, 1, 2, 2]):       # a a _ a b b  -> a a b
        x[t, 0, c] = 3.0
    seqs, nsl = ctc_ref.greedy_decode(x, [6])
    assert seqs == [[1, 1, 2]] and abs(nsl[0] + 18.0) < 1e-6
    assert ctc_ref.greedy_decode(x, [6], merge_repeated=False)[0] == [[1, 1, 1, 2, 2]]
    assert ctc_ref.greedy_decode(np.zeros((4, 1, V), np.float32), [4])[0] == [[0]]     # ties -> lowest index
    allblank = np.zeros((3, 1, V), np.float32)
    allblank[:, 0, 4] = 1.0
    assert ctc_ref.greedy_decode(allblank, [3])[0] == [[]]
    assert ctc_ref.greedy_decode(x, [2])[0] == [[1]]                                     # t >= len ignored
    assert ctc_ref.densify([[1, 2], []], 0).tolist() == [[1, 2], [0, 0]]
    assert ctc_ref.densify([[1, 2], []], -1).tolist() == [[1, 2], [-1, -1]]
