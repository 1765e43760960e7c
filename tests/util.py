"""Shared helpers of the test-suite."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def pkg():
    import podsdf_loader
    return podsdf_loader.load()


def rel_err(a, b):
    """max |a-b| relative to the RMS of the reference field b (the tolerance the north star states:
    1e-10 relative, FP64)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.sqrt(np.mean(b[np.isfinite(b)] ** 2)) if np.isfinite(b).any() else 1.0
    scale = scale if scale > 0 else 1.0
    same_nan = np.array_equal(np.isnan(a), np.isnan(b))
    if not same_nan:
        return np.inf
    d = np.abs(a - b)
    d = d[np.isfinite(b)]
    return float(d.max() / scale) if d.size else 0.0


def fix_signs(V, ref):
    """Flip the columns of V so that they correlate positively with the columns of ref."""
    s = np.sign(np.sum(V * ref, axis=0))
    s[s == 0] = 1.0
    return V * s[None, :]


TOL = 1e-10    # north-star tolerance (relative, FP64)
