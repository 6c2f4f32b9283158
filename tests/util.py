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
