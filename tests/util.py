"""Shared helpers of the parity tests."""
import numpy as np


def feature_err(a, b):
    """BASELINE.md section 5 metric: max |a-b| / max(|b|, 1)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1.0)))


FEATURE_TOL = 1e-4       # north star: features within 1e-4 (fp32)
CTC_RTOL, CTC_ATOL = 1e-3, 1e-5   # north star: CTC loss / grad within 1e-3 relative


def assert_ctc_grad_close(grad, grad_ref, logits, input_len):
    """grad = softmax - occupancy.  "Within 1e-3 relative" is applied to the two
    terms the kernel produces (both are probabilities): |d| <= rtol * (|g_ref| +
    occupancy_ref) + atol.  A plain rtol on the difference would be meaningless
    where softmax and occupancy cancel."""
    grad = np.asarray(grad, dtype=np.float64)
    grad_ref = np.asarray(grad_ref, dtype=np.float64)
    x = np.asarray(logits, dtype=np.float64)
    m = x.max(-1, keepdims=True)
    y = np.exp(x - m)
    y /= y.sum(-1, keepdims=True)
    T, B, V = x.shape
    valid = np.arange(T)[:, None] < np.asarray(input_len)[None, :]
    y = y * valid[:, :, None]
    occ = np.abs(y - grad_ref)
    tol = CTC_RTOL * (np.abs(grad_ref) + occ) + CTC_ATOL
    bad = np.abs(grad - grad_ref) > tol
    assert not bad.any(), ("grad mismatch", int(bad.sum()), float(np.abs(grad - grad_ref)[bad].max()))


def zscore_feature_err(got, ref_z, ref_raw):
    """Error metric for the z-scored features: max |a-b| / max(|b|,1), with every
    column's tolerance widened by max(1, 0.05/std_col): the GPU path keeps the
    log-spectrogram in float32 (absolute error ~1e-6 at log|X| ~ 15) and the z-score
    divides that by the column's standard deviation, so columns that are nearly
    constant over time (std << 0.05) are ill-conditioned for ANY float32 pipeline."""
    got = np.asarray(got, dtype=np.float64)
    ref_z = np.asarray(ref_z, dtype=np.float64)
    if got.size == 0:
        return 0.0
    sd = np.asarray(ref_raw, dtype=np.float64).std(axis=0)
    cond = np.maximum(1.0, 0.05 / np.maximum(sd, 1e-300))
    cond[sd < 10 * np.finfo(np.float64).eps] = 1.0     # constant columns: both sides give 0
    return float(np.max(np.abs(got - ref_z) / np.maximum(np.abs(ref_z), 1.0) / cond[None, :]))


def zscore_err_report(got, ref_z, ref_raw):
    """(un-widened error, widened error, number of columns that needed the 0.05/std allowance, i.e. whose
    un-widened error exceeds FEATURE_TOL) for one utterance's z-scored features."""
    got = np.asarray(got, dtype=np.float64)
    ref_z = np.asarray(ref_z, dtype=np.float64)
    if got.size == 0:
        return 0.0, 0.0, 0
    e = np.abs(got - ref_z) / np.maximum(np.abs(ref_z), 1.0)
    col = e.max(axis=0)
    sd = np.asarray(ref_raw, dtype=np.float64).std(axis=0)
    cond = np.maximum(1.0, 0.05 / np.maximum(sd, 1e-300))
    cond[sd < 10 * np.finfo(np.float64).eps] = 1.0
    return float(col.max()), float((col / cond).max()), int((col > FEATURE_TOL).sum())


def record_parity(key, values):
    """Merge one entry into gpurun_out/parity.json (worst errors per config)."""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "parity.json")
    try:
        data = json.load(open(path))
    except Exception:
        data = {}
    data[key] = values
    json.dump(data, open(path, "w"), indent=1, sort_keys=True)
