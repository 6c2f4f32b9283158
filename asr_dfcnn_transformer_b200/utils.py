"""Drop-in for the helpers of the reference's ``util/utils.py`` that sit right after the
hot path, computed on the device:

  build_LFR_features(inputs, m, n)   utils.py:7-31   (stack m frames, skip n)
  edit_distance / label_error_rate   tf.edit_distance(decoded, labels) + reduce_mean,
                                     lm_and_am/model/acoustic_model2.py:72-73
  GetEditDistance(str1, str2)        utils.py:43-53  (host: difflib opcodes, not Levenshtein)
"""
import difflib

import numpy as np

from . import _lib


def lfr_batch(features, frame_offsets, m=4, n=3, stream=None):
    """LFR stacking of a ragged device batch.  features: float32 ``[sum T_b, D]`` device
    tensor, frame_offsets: host int64 ``[B+1]``.  Returns (out ``[sum ceil(T_b/n), m*D]``,
    out_offsets host int64 ``[B+1]``)."""
    torch = _lib.require_cuda()
    fo = np.asarray(frame_offsets, dtype=np.int64)
    T = np.diff(fo)
    out_rows = -(-T // n)
    oo = np.concatenate([[0], np.cumsum(out_rows)]).astype(np.int64)
    dev = features.device
    D = int(features.shape[1])
    out = torch.empty((int(oo[-1]), m * D), dtype=torch.float32, device=dev)
    fo_d = torch.from_numpy(fo).to(dev)
    oo_d = torch.from_numpy(oo).to(dev)
    st = _lib.lib().asrk_lfr_run(_lib.ptr(features), _lib.ptr(fo_d), _lib.ptr(out), _lib.ptr(oo_d), len(T), D,
                                 int(m), int(n), int(oo[-1]), _lib.stream_ptr(stream))
    _lib.check(st, "asrk_lfr_run")
    return out, oo


def build_LFR_features(inputs, m, n):
    """utils.py:7-31 -- ``inputs`` is a T x D array; returns ceil(T/n) x (m*D), same dtype."""
    torch = _lib.require_cuda()
    x = np.ascontiguousarray(np.asarray(inputs))
    if x.ndim != 2:
        raise ValueError("build_LFR_features expects a T x D array")
    if x.shape[0] == 0:
        raise ValueError("need at least one array to concatenate")       # np.vstack([]) in the reference
    t = torch.from_numpy(x.astype(np.float32)).cuda()
    out, _ = lfr_batch(t, [0, x.shape[0]], m, n)
    return out.cpu().numpy().astype(x.dtype if x.dtype.kind == "f" else np.float64)


def edit_distance(hyp, hyp_len, truth, truth_len, normalize=True, stream=None):
    """Per-utterance Levenshtein distance of the decoded tokens against the labels
    (tf.edit_distance; divided by the label length when ``normalize``).  hyp: int32
    ``[B, Hmax]`` device tensor (e.g. CtcResult.tokens), truth: int32 ``[B, <=64]``."""
    torch = _lib.require_cuda()
    dev = hyp.device
    truth = torch.as_tensor(np.asarray(truth, dtype=np.int32) if not torch.is_tensor(truth) else truth).to(dev)
    truth = truth.to(torch.int32).contiguous()
    hl = torch.as_tensor(np.asarray(hyp_len, dtype=np.int32) if not torch.is_tensor(hyp_len) else hyp_len).to(dev)
    tl = torch.as_tensor(np.asarray(truth_len, dtype=np.int32) if not torch.is_tensor(truth_len) else truth_len).to(dev)
    hyp = hyp.to(torch.int32).contiguous()
    B = int(hyp.shape[0])
    out = torch.empty(B, dtype=torch.float32, device=dev)
    st = _lib.lib().asrk_edit_distance_run(_lib.ptr(hyp), int(hyp.shape[1]) if hyp.dim() > 1 else 0,
                                           _lib.ptr(hl.to(torch.int32).contiguous()), _lib.ptr(truth),
                                           int(truth.shape[1]) if truth.dim() > 1 else 0,
                                           _lib.ptr(tl.to(torch.int32).contiguous()), B, 1 if normalize else 0,
                                           _lib.ptr(out), _lib.stream_ptr(stream))
    _lib.check(st, "asrk_edit_distance_run")
    return out


def label_error_rate(hyp, hyp_len, truth, truth_len):
    """tf.reduce_mean(tf.edit_distance(...)) -- acoustic_model2.py:72-73."""
    return edit_distance(hyp, hyp_len, truth, truth_len, normalize=True).mean()


def GetEditDistance(str1, str2):
    """utils.py:43-53: NOT a Levenshtein distance but the cost of difflib's opcode blocks -- a
    'replace' block costs its longer side, 'insert' the inserted length, 'delete' the deleted one."""
    cost = {"replace": lambda di, dj: max(di, dj), "insert": lambda di, dj: dj, "delete": lambda di, dj: di,
            "equal": lambda di, dj: 0}
    ops = difflib.SequenceMatcher(None, str1, str2).get_opcodes()
    return sum(cost[tag](i2 - i1, j2 - j1) for tag, i1, i2, j1, j2 in ops)
