"""Seeded datasets shared by the golden generator and the tests (regenerable anywhere)."""
import numpy as np


def gmm(n, d, k=64, seed=0, dtype=np.float32, return_labels=False):
    """Gaussian mixture of SURVEY.md section 8(d): centres N(0, 2^2), unit noise."""
    rng = np.random.default_rng(seed)
    centers = rng.normal(0, 2, (k, d))
    lab = rng.integers(0, k, n)
    X = (centers[lab] + rng.normal(0, 1, (n, d))).astype(dtype)
    return (X, lab) if return_labels else X


def digits():
    from sklearn.datasets import load_digits

    X, y = load_digits(return_X_y=True)
    return X, y


def rings3d(n=1000, seed=3):
    """Two interlocked noisy rings in 3-D (D <= 15: sklearn takes its kd-tree branch)."""
    rng = np.random.default_rng(seed)
    t = rng.uniform(0, 2 * np.pi, n)
    half = n // 2
    a = np.stack([np.cos(t[:half]), np.sin(t[:half]), np.zeros(half)], axis=1)
    b = np.stack([1 + np.cos(t[half:]), np.zeros(n - half), np.sin(t[half:])], axis=1)
    X = np.concatenate([a, b]) + rng.normal(0, 0.05, (n, 3))
    y = np.concatenate([np.zeros(half, dtype=np.int64), np.ones(n - half, dtype=np.int64)])
    p = rng.permutation(n)
    # float32-representable values so the device (fp32 master copy) sees identical inputs
    return X[p].astype(np.float32).astype(np.float64), y[p]


def blobs2d(n=2000, seed=5):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-4, 4, (6, 2))
    lab = rng.integers(0, 6, n)
    X = c[lab] + rng.normal(0, 0.4, (n, 2))
    return X.astype(np.float32).astype(np.float64), lab


def load(name):
    """Return (X, y) for a dataset key used in the golden files."""
    if name == "digits":
        return digits()
    if name == "rings3d":
        return rings3d()
    if name == "blobs2d":
        return blobs2d()
    if name.startswith("gmm:"):
        # gmm:n:d:k:seed:dtype
        _, n, d, k, seed, dt = name.split(":")
        X, lab = gmm(int(n), int(d), int(k), int(seed), np.dtype(dt), return_labels=True)
        return X, lab
    raise KeyError(name)
