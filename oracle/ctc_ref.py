"""CPU restatement (numpy, float64) of the CTC loss/gradient and greedy decode the
reference reaches through Keras/TensorFlow.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never imported by the
product package.  PARITY UNPINNED against the reference itself: the arithmetic
is in un-vendored third-party code and the reference holds no test vector for it.

Reference call sites
  * /root/reference/lm_and_am/model/cnn_ctc.py:149-152  ``K.ctc_batch_cost``
    (Keras 2.3.1 ``tensorflow_backend.ctc_batch_cost``: ``log(transpose(y_pred)
    + 1e-7)``, length-masked dense->sparse labels, ``tf.nn.ctc_loss`` with the
    blank = V-1, ``[B,1]`` output)
  * /root/reference/lm_and_am/model/acoustic_model2.py:68,79-80
    ``tf.nn.ctc_loss_v2(..., blank_index=V-1)`` on time-major logits, labels via
    ``dense_to_sparse`` (:71, which drops every 0 entry)
  * /root/reference/lm_and_am/model/acoustic_model2.py:69 and
    /root/reference/util/utils.py:57-66 ``tf.nn.ctc_greedy_decoder`` /
    ``K.ctc_decode(greedy=True)``

Published algorithm restated (TensorFlow 1.14 ``core/util/ctc/ctc_loss_calculator
.{h,cc}``, ``ctc_loss_util.h``, ``core/kernels/ctc_decoder_ops.cc``):
  y = softmax(logits) per frame; l' = [b, l1, b, ..., lL, b]; log-space alpha
  (includes y_t) and beta (excludes y_t); log p = LSE_u alpha(u,0)+beta(u,0);
  loss = -log p; d loss / d logits[t,v] = y[t,v] - exp(LSE_{u: l'_u = v}(alpha
  (u,t) + beta(u,t)) - log p) for t < len and 0 for t >= len.
"""
import numpy as np

NEG_INF = -np.inf


def _lse(*xs):
    m = max(xs)
    if m == NEG_INF:
        return NEG_INF
    return m + np.log(sum(np.exp(x - m) for x in xs))


def log_softmax(x):
    x = np.asarray(x, dtype=np.float64)
    m = x.max(axis=-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=-1, keepdims=True))


def ctc_loss_grad_single(logits, labels, blank):
    """One utterance.  logits float[T,V] (already cut to its length), labels
    int[L].  Returns (loss, grad[T,V], feasible)."""
    logits = np.asarray(logits, dtype=np.float64)
    T, V = logits.shape
    labels = [int(v) for v in labels]
    L = len(labels)
    lp = log_softmax(logits)
    y = np.exp(lp)
    lprime = [blank]
    for c in labels:
        lprime += [c, blank]
    U = len(lprime)

    alpha = np.full((T, U), NEG_INF)
    beta = np.full((T, U), NEG_INF)
    # ctc_loss_calculator.cc CalculateForwardVariables
    alpha[0, 0] = lp[0, blank]
    if U > 1:
        alpha[0, 1] = lp[0, lprime[1]]
    for t in range(1, T):
        lo = max(0, U - 2 * (T - t))
        hi = min(U, 2 * (t + 1))
        for u in range(lo, hi):
            s = alpha[t - 1, u]
            if u >= 1:
                s = _lse(s, alpha[t - 1, u - 1])
            if u >= 2 and lprime[u] != blank and lprime[u] != lprime[u - 2]:
                s = _lse(s, alpha[t - 1, u - 2])
            alpha[t, u] = s + lp[t, lprime[u]]
    # CalculateBackwardVariables (beta excludes y_t)
    beta[T - 1, U - 1] = 0.0
    if U > 1:
        beta[T - 1, U - 2] = 0.0
    for t in range(T - 2, -1, -1):
        lo = max(0, U - 2 * (T - t))
        hi = min(U, 2 * (t + 1))
        for u in range(lo, hi):
            s = beta[t + 1, u] + lp[t + 1, lprime[u]]
            if u + 1 < U:
                s = _lse(s, beta[t + 1, u + 1] + lp[t + 1, lprime[u + 1]])
            if u + 2 < U and lprime[u] != blank and lprime[u] != lprime[u + 2]:
                s = _lse(s, beta[t + 1, u + 2] + lp[t + 1, lprime[u + 2]])
            beta[t, u] = s
    log_p = NEG_INF
    for u in range(U):
        log_p = _lse(log_p, alpha[0, u] + beta[0, u])
    if log_p == NEG_INF:
        # no valid path: TF returns +inf loss and dy = y
        return np.inf, y.copy(), False
    grad = y.copy()
    for t in range(T):
        acc = {}
        for u in range(U):
            ab = alpha[t, u] + beta[t, u]
            if ab == NEG_INF:
                continue
            c = lprime[u]
            acc[c] = _lse(acc.get(c, NEG_INF), ab)
        for c, v in acc.items():
            grad[t, c] -= np.exp(v - log_p)
    return -log_p, grad, True


def ctc_loss_grad_batch(logits_tbv, labels, label_len, input_len, blank, label_mode="by_length",
                        impl="vec"):
    """Batch, TF layout: logits float[T,B,V] time-major, labels int[B,Lmax].

    label_mode 'by_length' = Keras ``ctc_label_dense_to_sparse`` (keeps zeros,
    masks by label_len); 'drop_zeros' = ``tf.contrib.layers.dense_to_sparse``
    (acoustic_model2.py:71: every 0 entry is dropped, label_len ignored).
    Returns loss[B] float64, grad[T,B,V] float64 (zero for t >= input_len)."""
    logits_tbv = np.asarray(logits_tbv)
    T, B, V = logits_tbv.shape
    loss = np.zeros(B)
    grad = np.zeros((T, B, V))
    feasible = np.ones(B, dtype=bool)
    for b in range(B):
        if label_mode == "by_length":
            lab = list(labels[b][: int(label_len[b])])
        else:
            lab = [int(v) for v in labels[b] if int(v) != 0]
        tl = int(input_len[b])
        fn = ctc_loss_grad_single_vec if impl == "vec" else ctc_loss_grad_single
        l, g, ok = fn(logits_tbv[:tl, b, :], lab, blank)
        loss[b] = l
        grad[:tl, b, :] = g
        feasible[b] = ok
    return loss, grad, feasible


def keras_ctc_batch_cost(y_true, y_pred, input_length, label_length):
    """Keras 2.3.1 ``K.ctc_batch_cost`` as called at cnn_ctc.py:149-152.
    y_pred = softmax output [B,T,V]; returns (loss[B,1], d loss / d y_pred)."""
    y_pred = np.asarray(y_pred, dtype=np.float64)
    B, T, V = y_pred.shape
    eps = 1e-7
    x = np.log(np.transpose(y_pred, (1, 0, 2)) + eps)
    il = np.asarray(input_length).reshape(-1).astype(np.int64)
    ll = np.asarray(label_length).reshape(-1).astype(np.int64)
    loss, gx, _ = ctc_loss_grad_batch(x, np.asarray(y_true).astype(np.int64), ll, il, V - 1)
    # chain rule through x = log(p + eps)
    gp = np.transpose(gx, (1, 0, 2)) / (y_pred + eps)
    return loss.reshape(B, 1), gp


def greedy_decode(logits_tbv, input_len, blank=None, merge_repeated=True):
    """TF ``CTCGreedyDecoderOp`` (ctc_decoder_ops.cc): per frame the FIRST
    maximum (strict '>' scan), emit when != blank and (not merge or != prev);
    prev is updated on every frame; neg_sum_logits = -sum_t max_v logits.
    Returns (list of int lists, neg_sum_logits[B])."""
    logits_tbv = np.asarray(logits_tbv)
    T, B, V = logits_tbv.shape
    if blank is None:
        blank = V - 1
    out, nsl = [], np.zeros(B, dtype=np.float64)
    for b in range(B):
        prev = -1
        seq = []
        for t in range(int(input_len[b])):
            row = logits_tbv[t, b]
            c = int(np.argmax(row))  # first maximum
            nsl[b] += -float(row[c])
            if c != blank and not (merge_repeated and c == prev):
                seq.append(c)
            prev = c
        out.append(seq)
    return out, nsl


def densify(seqs, pad, width=None):
    """``tf.sparse_tensor_to_dense(default_value=pad)`` (test.py:51 pads with 0;
    Keras ``ctc_decode`` pads with -1)."""
    w = max([len(s) for s in seqs] + [0]) if width is None else width
    out = np.full((len(seqs), w), pad, dtype=np.int64)
    for i, s in enumerate(seqs):
        out[i, : len(s)] = s
    return out


def ctc_loss_grad_single_vec(logits, labels, blank):
    """Same algorithm as ``ctc_loss_grad_single`` with the loop over lattice
    states vectorised (numpy) -- for the long-utterance configurations where the
    scalar loops would take minutes.  Checked against the scalar version in
    tests/test_oracle_ctc.py."""
    logits = np.asarray(logits, dtype=np.float64)
    T, V = logits.shape
    labels = np.asarray(labels, dtype=np.int64).reshape(-1)
    L = len(labels)
    lp = log_softmax(logits)
    U = 2 * L + 1
    lprime = np.full(U, blank, dtype=np.int64)
    lprime[1::2] = labels
    skip = np.zeros(U, dtype=bool)
    skip[2:] = (lprime[2:] != blank) & (lprime[2:] != lprime[:-2])
    skip_b = np.zeros(U, dtype=bool)
    skip_b[:-2] = (lprime[:-2] != blank) & (lprime[:-2] != lprime[2:])
    lpl = lp[:, lprime]                      # [T,U] log y_t(l'_u)
    uidx = np.arange(U)
    alpha = np.full((T, U), NEG_INF)
    beta = np.full((T, U), NEG_INF)
    alpha[0, 0] = lpl[0, 0]
    if U > 1:
        alpha[0, 1] = lpl[0, 1]
    with np.errstate(invalid="ignore"):
        for t in range(1, T):
            a = alpha[t - 1]
            s = a.copy()
            s[1:] = np.logaddexp(s[1:], a[:-1])
            s2 = np.full(U, NEG_INF)
            s2[2:] = a[:-2]
            s = np.where(skip, np.logaddexp(s, s2), s)
            s = s + lpl[t]
            lo = max(0, U - 2 * (T - t))
            hi = min(U, 2 * (t + 1))
            s[(uidx < lo) | (uidx >= hi)] = NEG_INF
            alpha[t] = s
        beta[T - 1, U - 1] = 0.0
        if U > 1:
            beta[T - 1, U - 2] = 0.0
        for t in range(T - 2, -1, -1):
            bn = beta[t + 1] + lpl[t + 1]
            s = bn.copy()
            s[:-1] = np.logaddexp(s[:-1], bn[1:])
            s2 = np.full(U, NEG_INF)
            s2[:-2] = bn[2:]
            s = np.where(skip_b, np.logaddexp(s, s2), s)
            lo = max(0, U - 2 * (T - t))
            hi = min(U, 2 * (t + 1))
            s[(uidx < lo) | (uidx >= hi)] = NEG_INF
            beta[t] = s
        ab0 = alpha[0] + beta[0]
        m = ab0.max()
        if m == NEG_INF:
            return np.inf, np.exp(lp), False
        log_p = m + np.log(np.exp(ab0 - m).sum())
        occ = np.exp(alpha + beta - log_p)   # [T,U]; exp(-inf) = 0
    grad = np.exp(lp)
    sub = np.zeros((T, V))
    for u in range(U):                       # duplicates accumulate
        sub[:, lprime[u]] += occ[:, u]
    grad -= sub
    return -log_p, grad, True
