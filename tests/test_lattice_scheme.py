"""CPU: the numerical scheme of the fused CTC kernel's lattice (csrc/ctc.cu, phase B), emulated in
numpy float32 -- log2-domain alpha / beta (beta excludes y_t, TF convention), every column stored
relative to a level that has the PREVIOUS stored column's maximum subtracted, levels accumulated in
float64 -- against the float64 oracle.  Pins the scheme itself (no GPU needed): no drift with T and
no loss of low-mass terminal states."""
import numpy as np
import pytest

from oracle import ctc_ref, synth

F = np.float32
NEG = F(-np.inf)


def _lse(*xs):
    m = xs[0]
    for x in xs[1:]:
        m = np.maximum(m, x)
    ms = np.where(np.isneginf(m), F(0), m)
    s = sum(np.exp2((x - ms).astype(F)).astype(F) for x in xs)
    with np.errstate(divide="ignore"):
        return (m + np.log2(s).astype(F)).astype(F)


def lattice_fp32(logits, labels, blank):
    """Returns (loss, occupancy[T, 2L+1]) with the kernel's arithmetic."""
    x = np.asarray(logits, dtype=F)
    T, V = x.shape
    L = len(labels)
    m = x.max(-1, keepdims=True)
    lse = (m[:, 0] + np.log(np.exp(x - m).sum(-1, dtype=F))).astype(F)
    ext = [blank]
    for c in labels:
        ext += [int(c), blank]
    U = len(ext)
    ly = ((x[:, ext] - lse[:, None]) * F(1.4426950408889634)).astype(F)        # log2 y_t(l'_u)
    skip = np.zeros(U, bool)
    for u in range(3, U, 2):
        skip[u] = ext[u] != ext[u - 2]
    al = np.full((T, U), NEG)
    C = np.zeros(T)
    a = np.full(U, NEG)
    a[0] = ly[0, 0]
    if U > 1:
        a[1] = ly[0, 1]
    mp, lvl = F(0), 0.0
    for t in range(T):
        if t > 0:
            p1 = np.r_[NEG, a][:U].astype(F)
            p2 = np.r_[NEG, NEG, a][:U].astype(F)
            p2 = np.where(skip, p2, NEG)
            a = (ly[t] + _lse(a, p1, p2) - mp).astype(F)
            lvl += float(mp)
        al[t], C[t] = a, lvl
        c = a.max()
        mp = F(0) if np.isneginf(c) else c
    be = np.full((T, U), NEG)
    D = np.zeros(T)
    b = np.full(U, NEG)
    b[U - 1] = F(0)
    if U > 1:
        b[U - 2] = F(0)
    skipn = np.r_[skip, False, False][2:U + 2]                                  # u -> u + 2 allowed
    mp, lvl = F(0), 0.0
    for t in range(T - 1, -1, -1):
        if t < T - 1:
            e = (b + ly[t + 1]).astype(F)
            n1 = np.r_[e, NEG][1:U + 1].astype(F)
            n2 = np.where(skipn, np.r_[e, NEG, NEG][2:U + 2].astype(F), NEG)
            b = (_lse(e, n1, n2) - mp).astype(F)
            lvl += float(mp)
        be[t], D[t] = b, lvl
        d = b.max()
        mp = F(0) if np.isneginf(d) else d
    fin = _lse(al[T - 1, U - 1:U], al[T - 1, U - 2:U - 1] if U > 1 else np.array([NEG]))[0]
    logp2 = C[T - 1] + float(fin)
    K = (C + D - logp2).astype(F)
    occ = np.exp2((al + be + K[:, None]).astype(F))
    return -logp2 * np.log(2.0), occ


@pytest.mark.parametrize("T,L,scale", [(30, 8, 3.0), (88, 24, 3.0), (200, 31, 6.0), (64, 0, 3.0)])
def test_scheme_matches_the_float64_oracle(T, L, scale):
    rng = np.random.default_rng(T + L)
    x = synth.logits_tbv(rng, T, 1, 1424, scale)[:, 0]
    labels = list(synth.labels_with_repeats(rng, L, 1424))
    while L and L + synth.n_repeats(labels) > T:
        labels = labels[:-1]
    loss, occ = lattice_fp32(x, labels, 1423)
    ref_loss, ref_grad, ok = ctc_ref.ctc_loss_grad_single(x, labels, 1423)
    assert ok and abs(loss - ref_loss) <= 1e-3 * max(1.0, abs(ref_loss))
    # occupancy summed per class = softmax - gradient of the oracle
    xx = np.asarray(x, np.float64)
    y = np.exp(xx - xx.max(-1, keepdims=True))
    y /= y.sum(-1, keepdims=True)
    ext = [1423]
    for c in labels:
        ext += [int(c), 1423]
    got = np.zeros_like(y)
    for u, c in enumerate(ext):
        got[:, c] += occ[:, u]
    assert np.abs(got - (y - ref_grad)).max() <= 1e-3
    # the states of a frame share all the mass; fp32 rounding of ~1e-7 per step adds up over T frames
    assert np.abs(occ.sum(-1) - 1.0).max() <= 2.5e-6 * T


def test_low_mass_terminal_states_survive():
    """Frames whose distribution makes the terminal states carry ~2^-200 of their column: a scheme that
    rescales in the LINEAR domain by the column maximum loses them (denormals); this one must not."""
    rng = np.random.default_rng(9)
    T, V, labels = 40, 64, [5, 9, 5, 7]
    x = rng.standard_normal((T, V)).astype(np.float32)
    x[:, 63] += 12.0                                          # blank dominates early ...
    x[T - 6:, 63] -= 30.0                                     # ... and is nearly impossible at the end
    loss, occ = lattice_fp32(x, labels, 63)
    ref_loss, _, ok = ctc_ref.ctc_loss_grad_single(x, labels, 63)
    assert ok and abs(loss - ref_loss) <= 1e-3 * abs(ref_loss)
    assert np.abs(occ.sum(-1) - 1.0).max() <= 1e-4
