"""ORACLE (test infrastructure only): restatement of the helpers next to the hot path.

  build_LFR_features   /root/reference/util/utils.py:7-31 (restated as an index gather)
  edit_distance        tf.edit_distance semantics (tensorflow/core/kernels/edit_distance_op.cc,
                       lib/gtl/edit_distance.h: Levenshtein; normalize divides by the truth
                       length; empty truth -> inf for a non-empty hypothesis, 0 otherwise)
                       as called at lm_and_am/model/acoustic_model2.py:72.  TensorFlow is not
                       installable here: parity unpinned by the reference, known answers below
                       are hand-computed.
"""
import numpy as np


def build_LFR_features(inputs, m, n):
    """Restated as one gather: output row i stacks the input rows i*n .. i*n+m-1, and a row index
    past the end is replaced by the LAST row (utils.py:20-30: the tail is padded by repeating
    inputs[-1]); ceil(T/n) output rows (utils.py:21)."""
    inputs = np.asarray(inputs)
    T = inputs.shape[0]
    rows = -(-T // n)
    src = np.minimum(np.arange(rows)[:, None] * n + np.arange(m)[None, :], T - 1)     # [rows, m]
    return inputs[src].reshape(rows, m * inputs.shape[1])


def levenshtein(hyp, truth):
    hyp, truth = list(hyp), list(truth)
    prev = list(range(len(truth) + 1))
    for i, h in enumerate(hyp, 1):
        cur = [i]
        for j, t in enumerate(truth, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (0 if h == t else 1)))
        prev = cur
    return prev[-1]


def edit_distance(hyps, truths, normalize=True):
    out = []
    for h, t in zip(hyps, truths):
        d = float(levenshtein(h, t))
        if normalize:
            d = d / len(t) if len(t) else (float("inf") if len(h) else 0.0)
        out.append(d)
    return np.array(out, dtype=np.float64)
