"""Small host-side formulas shared by the engine implementations."""

from __future__ import annotations

import numpy as np


def class_entropy(hist: np.ndarray) -> np.ndarray:
    """Per-neuron base-2 entropy of the class histogram [M, C].

    Reference: `scipy.stats.entropy(np.bincount(y[winners == j]), base=2)` per neuron
    (dbgsom/BaseSom.py:547-551); a neuron without samples gets 0.0 there (entropy of an
    empty count vector), not NaN.
    """
    import scipy.stats

    hist = np.asarray(hist, dtype=np.float64)
    out = np.zeros(hist.shape[0])
    live = hist.sum(axis=1) > 0
    if live.any():
        out[live] = scipy.stats.entropy(hist[live], base=2, axis=1)
    return out


def column_moments_to_stats(n: int, shift_sum: np.ndarray, shift_sumsq: np.ndarray) -> dict:
    """Turn per-column sums of (x - c) and (x - c)^2 into the two scalars `fit` needs.

    total_variance = sum_d var(X[:, d]) with ddof=0      (dbgsom/BaseSom.py:363)
    std_norm       = || std(X, axis=0, ddof=1) ||_2      (dbgsom/BaseSom.py:380-383)
    """
    ss = np.maximum(shift_sumsq - shift_sum * shift_sum / n, 0.0)
    return {
        "n_samples": int(n),
        "total_variance": float((ss / n).sum()),
        "std_norm": float(np.sqrt((ss / (n - 1)).sum())) if n > 1 else float("nan"),
    }
