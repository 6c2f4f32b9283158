"""CTC pins that do NOT come from the restated alpha/beta recursion (oracle/ctc_ref.py):

1. cases whose loss and FULL gradient are derived on paper (docstrings below);
2. the definition of CTC itself (Graves et al. 2006, eq. 3): p(l|x) = sum over all alignments pi in V^T with
   collapse(pi) = l of prod_t y_t(pi_t), by brute-force enumeration of every alignment, and its gradient
   d(-log p)/dx[t,v] = y_t(v) - (1/p) sum_{pi: pi_t = v, collapse(pi) = l} P(pi)  (softmax inside the op, as in
   TensorFlow's CTCLossOp) -- no dynamic programme involved.

The CPU tests hold the oracle to these; the GPU tests hold the CUDA kernels to them (both input kinds: logits,
and the probabilities of the Keras path with log(p + 1e-7) formed inside the kernel).
"""
import itertools

import numpy as np
import pytest

from oracle import ctc_ref


def softmax(x):
    e = np.exp(x - x.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def collapse(path, blank):
    out, prev = [], None
    for c in path:
        if c != prev and c != blank:
            out.append(c)
        prev = c
    return out


def brute_force(x, label, blank):
    """x [T,V] logits of ONE utterance.  Returns (loss, grad [T,V]) from the definition."""
    T, V = x.shape
    y = softmax(x.astype(np.float64))
    p = 0.0
    occ = np.zeros((T, V))
    for path in itertools.product(range(V), repeat=T):
        if collapse(path, blank) != list(label):
            continue
        pp = 1.0
        for t, c in enumerate(path):
            pp *= y[t, c]
        p += pp
        for t, c in enumerate(path):
            occ[t, c] += pp
    return -np.log(p), y - occ / p


def paper_case_single_path():
    """V = {a=0, b=1, blank=2}, T = 3, label "a a".  A repeated label needs a blank between its two
    occurrences, so with T = L + repeats = 3 frames there is exactly ONE alignment: (a, blank, a).
        p    = y0(a) y1(blank) y2(a)                      loss = -log p
        occupancy[t, v] = 1 for the class the alignment emits at t, else 0
        dL/dx[t, :] = y_t - onehot(class at t)            (softmax inside the op)
    With x = log of the rows below (so y = the rows):  p = 0.5 * 0.25 * 0.2 = 0.025, loss = log 40."""
    y = np.array([[0.5, 0.3, 0.2], [0.5, 0.25, 0.25], [0.2, 0.7, 0.1]])
    x = np.log(y)
    loss = np.log(40.0)
    grad = y - np.array([[1, 0, 0], [0, 0, 1], [1, 0, 0]], dtype=np.float64)
    return x, [0, 0], 2, loss, grad


def paper_case_uniform():
    """V = {a=0, b=1, blank=2}, T = 2, label "a", all logits 0 (y = 1/3 everywhere).  Alignments that
    collapse to "a": (a,a), (a,blank), (blank,a), each with probability 1/9:
        p = 3/9 = 1/3, loss = log 3
        frame 0: a is emitted by (a,a) and (a,blank): occupancy 2/3; blank by (blank,a): 1/3
        frame 1: a by (a,a) and (blank,a): 2/3; blank by (a,blank): 1/3
        dL/dx = y - occupancy = [[1/3 - 2/3, 1/3 - 0, 1/3 - 1/3]] * 2 = [[-1/3, 1/3, 0], [-1/3, 1/3, 0]]."""
    x = np.zeros((2, 3))
    grad = np.array([[-1 / 3, 1 / 3, 0.0], [-1 / 3, 1 / 3, 0.0]])
    return x, [0], 2, np.log(3.0), grad


def enumerated_cases():
    rng = np.random.default_rng(77)
    cases = []
    # (T, V, label, blank): repeats, a label equal to class 0, empty label, T barely enough, blank not last
    for T, V, label, blank in [(4, 3, [0, 0], 2), (5, 4, [1, 1, 2], 3), (5, 3, [0, 1, 0], 2), (3, 4, [], 3),
                               (6, 3, [1, 0], 2), (4, 4, [2, 0, 2], 3), (5, 3, [1, 2], 0), (1, 3, [1], 2)]:
        x = 2.0 * rng.standard_normal((T, V))
        cases.append((x, label, blank))
    return cases


def _oracle(x, label, blank, T_total=None):
    T, V = x.shape
    Tt = T if T_total is None else T_total
    xx = np.zeros((Tt, 1, V))
    xx[:T, 0] = x
    lab = np.zeros((1, max(len(label), 1)), dtype=np.int32)
    lab[0, :len(label)] = label
    loss, grad, ok = ctc_ref.ctc_loss_grad_batch(xx, lab, np.array([len(label)], np.int32), np.array([T], np.int32), blank)
    return float(loss[0]), grad[:, 0]


def test_oracle_matches_paper_and_definition():
    for x, label, blank, loss, grad in (paper_case_single_path(), paper_case_uniform()):
        lo, go = _oracle(x, label, blank)
        assert abs(lo - loss) < 1e-12 and np.abs(go - grad).max() < 1e-12
    for x, label, blank in enumerated_cases():
        lb, gb = brute_force(x, label, blank)
        lo, go = _oracle(x, label, blank, T_total=x.shape[0] + 2)      # two padding frames: zero gradient rows
        assert abs(lo - lb) < 1e-10, (label, lo, lb)
        assert np.abs(go[:x.shape[0]] - gb).max() < 1e-10
        assert not go[x.shape[0]:].any()


def _gpu_case(x, label, blank, kind):
    """Run ONE utterance through the C ABI (batch of 3 copies with different paddings, V padded to a multiple of 4
    with very negative logits / zero probabilities so that the vector path is taken too)."""
    import torch
    from asr_dfcnn_transformer_b200 import ctc
    T, V = x.shape
    outs = []
    for Vp in (V, (V + 3) // 4 * 4 + 4):
        xx = np.full((T + 2, 3, Vp), -60.0)
        xx[:T, :, :V] = x[:, None, :]
        xx[T:, :, :V] = 0.3                       # padding frames hold arbitrary values
        if blank == V - 1 and Vp != V:             # keep the blank the last class of the widened vocabulary
            xx[:, :, [V - 1, Vp - 1]] = xx[:, :, [Vp - 1, V - 1]]
            bl = Vp - 1
        else:
            bl = blank
        lab = np.zeros((3, max(len(label), 1) + 1), dtype=np.int32)
        lab[:, :len(label)] = label
        ll = np.full(3, len(label), np.int32)
        il = np.full(3, T, np.int32)
        if kind == "prob":
            inp = softmax(xx)                      # the kernel forms log(p + 1e-7) itself
        else:
            inp = xx
        t = torch.as_tensor(inp.astype(np.float32)).cuda()
        r = ctc.ctc_loss_grad(t, lab, ll, il, bl, input_kind="prob" if kind == "prob" else "logits")
        loss = r.loss.cpu().numpy().astype(np.float64)
        grad = r.grad.cpu().numpy().astype(np.float64)
        assert (r.row_status.cpu().numpy() == 0).all()
        assert not grad[T:].any()
        g = grad[:T, 1]
        if bl != blank:
            g[:, [V - 1, Vp - 1]] = g[:, [Vp - 1, V - 1]]
        if kind == "prob":
            # gradient w.r.t. p -> gradient w.r.t. x = log(p + eps):  dL/dx = dL/dp (p + eps)
            pp = inp[:T, 1].astype(np.float32).astype(np.float64)
            if bl != blank:
                pp[:, [V - 1, Vp - 1]] = pp[:, [Vp - 1, V - 1]]
            g = g * (pp + 1e-7)
        outs.append((float(loss[1]), g[:, :V], Vp))
    return outs


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["logits", "prob"])
def test_gpu_matches_paper_and_definition(kind):
    cases = [(x, label, blank, loss, grad) for x, label, blank, loss, grad in
             (paper_case_single_path(), paper_case_uniform())]
    for x, label, blank in enumerated_cases():
        lb, gb = brute_force(x, label, blank)
        cases.append((x, label, blank, lb, gb))
    for x, label, blank, loss, grad in cases:
        for lo, go, Vp in _gpu_case(x, label, blank, kind):
            # softmax over log(p + 1e-7) differs from softmax over x by O(V 1e-7) in the probabilities
            tol = 1e-3
            assert abs(lo - loss) <= tol * max(abs(loss), 1.0), (kind, label, Vp, lo, loss)
            assert np.abs(go - grad).max() <= tol, (kind, label, Vp, float(np.abs(go - grad).max()))
