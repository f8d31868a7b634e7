"""CPU oracle for the DBGSOM batch-SOM training epoch.  TEST INFRASTRUCTURE ONLY.

This module is a plain numpy (float64) restatement of the reference's hot path
(SandroMartens/DBGSOM, `dbgsom/BaseSom.py`).  It exists to *check* the CUDA path;
nothing under `dbgsom_b200/` may import it.  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s cpu_baseline / `--impl reference` legs use it.

Parity status: **pinned by generated goldens**.  The reference's own test-suite pins
nothing on this path, so the oracle is pinned against outputs of the
reference itself, produced in the build container by `tests/golden/make_golden.py`
(which imports the unmodified reference from /root/reference through `oracle/_refshim.py`)
and committed under `tests/golden/`.  `tests/test_oracle_golden.py` replays them.

Third-party arithmetic.  The reference's BMU search is not in its own source: it is
`sklearn.neighbors.NearestNeighbors(n_neighbors=k).fit(W).kneighbors(X)` (BaseSom.py:455-457),
scikit-learn 1.9.0 in this image (requirements.txt:4 only says >=1.2).  `bmu()` below calls
that same library entry point (it is installed on the GPU box as well); `bmu_expansion()`
restates the published algorithm of its brute-force branch (`EuclideanArgKmin64`:
||x||^2 - 2 x.w + ||w||^2 in float64, clamp at 0, strict-< heap so the lowest index wins
ties, sqrt at the end); for D <= 15 sklearn switches to an exact kd-tree instead, whose
distances differ from the expansion by its ~1e-7 absolute cancellation noise.

Every function cites the reference lines it follows (paths relative to the reference root).
"""

from __future__ import annotations

from math import exp, log, sqrt

import numpy as np

__all__ = [
    "total_variance",
    "growing_threshold",
    "initial_rows",
    "bmu",
    "bmu_expansion",
    "sample_weights",
    "voronoi_sums",
    "voronoi_centers",
    "current_sigma",
    "neighborhood",
    "smooth",
    "epoch_step",
    "hop_matrix_grid",
]


# --------------------------------------------------------------------------- init
def total_variance(X: np.ndarray) -> float:
    """`self._total_variance = np.var(data, axis=0).sum()` -- dbgsom/BaseSom.py:363.

    Evaluated in the dtype of X like the reference (float32 X -> float32 accumulation); the
    device path always reduces in float64, which differs by ~1e-7 relative for float32 X.
    """
    return np.var(X, axis=0).sum()


def growing_threshold(
    X: np.ndarray,
    spreading_factor: float = 0.5,
    threshold_method: str = "se",
    growth_criterion: str = "quantization_error",
) -> float:
    """`_calculate_growing_threshold` -- dbgsom/BaseSom.py:371-385."""
    if growth_criterion == "entropy":
        return spreading_factor
    if threshold_method == "classical":
        return -X.shape[1] * log(spreading_factor)
    std = np.std(X, axis=0, ddof=1)
    return float(150 * -log(spreading_factor) * np.linalg.norm(std))


def initial_rows(n_samples: int, random_state) -> np.ndarray:
    """Indices of the 4 start prototypes.

    The reference draws `default_rng(seed=random_state).choice(a=data, size=4,
    replace=False)` (dbgsom/BaseSom.py:423-424); choosing rows of a 2-D array equals
    choosing row indices with the same generator state.
    """
    rng = np.random.default_rng(seed=random_state)
    return rng.choice(n_samples, size=4, replace=False)


# --------------------------------------------------------------------------- BMU
def bmu(X: np.ndarray, W: np.ndarray, n_bmu: int = 1):
    """`_get_winning_neurons` -- dbgsom/BaseSom.py:446-464 (sklearn does the work)."""
    from sklearn.neighbors import NearestNeighbors

    nn = NearestNeighbors(n_neighbors=n_bmu)
    nn.fit(W)
    dist, idx = nn.kneighbors(X)
    idx = idx.T[0:n_bmu].T
    if n_bmu == 1:
        idx = idx.reshape(-1)
        dist = dist.reshape(-1)
    return dist, idx


def bmu_expansion(X: np.ndarray, W: np.ndarray, n_bmu: int = 1, chunk: int = 4096):
    """Restatement of sklearn's brute-force `EuclideanArgKmin64`.

    sklearn/metrics/_pairwise_distances_reduction/_argkmin.pyx.tp:471-510 (GEMM
    expansion per chunk), :285-295 (clamp + sqrt), sklearn/utils/_heap.pyx:46 (strict <:
    the lowest index wins exact ties).  Returns (dist [N] or [N,k], idx) ascending.
    """
    X = np.asarray(X, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    n = X.shape[0]
    wn = np.einsum("ij,ij->i", W, W)
    dist = np.empty((n, n_bmu))
    idx = np.empty((n, n_bmu), dtype=np.int64)
    for s in range(0, n, chunk):
        xs = X[s : s + chunk]
        d2 = np.einsum("ij,ij->i", xs, xs)[:, None] - 2.0 * (xs @ W.T) + wn[None, :]
        np.maximum(d2, 0.0, out=d2)
        # stable argsort == ascending distance, lowest index first on exact ties
        order = np.argsort(d2, axis=1, kind="stable")[:, :n_bmu]
        idx[s : s + chunk] = order
        dist[s : s + chunk] = np.sqrt(np.take_along_axis(d2, order, axis=1))
    if n_bmu == 1:
        return dist.reshape(-1), idx.reshape(-1)
    return dist, idx


def expansion_distance(X: np.ndarray, W: np.ndarray, idx: np.ndarray) -> np.ndarray:
    """Distance of sample i to prototype idx[i] in the arithmetic of sklearn's brute-force search
    (`sqrt(max(0, ||x||^2 - 2 x.w + ||w||^2))`, see `bmu_expansion`)."""
    X = np.asarray(X, dtype=np.float64)
    Wi = np.asarray(W, dtype=np.float64)[idx]
    d2 = np.einsum("ij,ij->i", X, X) - 2.0 * np.einsum("ij,ij->i", X, Wi) + np.einsum("ij,ij->i", Wi, Wi)
    return np.sqrt(np.maximum(d2, 0.0))


def bmu_with_gap(X: np.ndarray, W: np.ndarray, chunk: int = 4096):
    """(dist, idx, relative gap) of `bmu_expansion` and `relative_gap` from ONE pass over the distances."""
    X = np.asarray(X, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    n = X.shape[0]
    wn = np.einsum("ij,ij->i", W, W)
    dist, idx, gap = np.empty(n), np.empty(n, dtype=np.int64), np.empty(n)
    for s in range(0, n, chunk):
        xs = X[s : s + chunk]
        d2 = np.einsum("ij,ij->i", xs, xs)[:, None] - 2.0 * (xs @ W.T) + wn[None, :]
        np.maximum(d2, 0.0, out=d2)
        best = np.argmin(d2, axis=1)  # first minimum == lowest index on exact ties
        rows = np.arange(d2.shape[0])
        b = d2[rows, best]
        idx[s : s + chunk], dist[s : s + chunk] = best, np.sqrt(b)
        if d2.shape[1] < 2:
            gap[s : s + chunk] = np.inf
            continue
        d2[rows, best] = np.inf
        second = d2.min(axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            g = (second - b) / b
        g[b == 0] = np.where(second[b == 0] > 0, np.inf, 0.0)
        gap[s : s + chunk] = g
    return dist, idx, gap


def sqdist_exact(X: np.ndarray, W: np.ndarray, chunk: int = 2048) -> np.ndarray:
    """All squared distances by direct differences in float64 (no cancellation)."""
    X = np.asarray(X, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    out = np.empty((X.shape[0], W.shape[0]))
    for s in range(0, X.shape[0], chunk):
        diff = X[s : s + chunk, None, :] - W[None, :, :]
        out[s : s + chunk] = np.einsum("nmd,nmd->nm", diff, diff)
    return out


def relative_gap(X: np.ndarray, W: np.ndarray, chunk: int = 4096) -> np.ndarray:
    """(d2_second - d2_best) / d2_best per sample in float64 (expansion form).

    This is the quantity the parity gate of BASELINE.json is stated on: BMU indices must
    match wherever it is >= 1e-6.
    """
    X = np.asarray(X, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    wn = np.einsum("ij,ij->i", W, W)
    gap = np.empty(X.shape[0])
    for s in range(0, X.shape[0], chunk):
        xs = X[s : s + chunk]
        d2 = np.einsum("ij,ij->i", xs, xs)[:, None] - 2.0 * (xs @ W.T) + wn[None, :]
        np.maximum(d2, 0.0, out=d2)
        if d2.shape[1] < 2:
            gap[s : s + chunk] = np.inf
            continue
        part = np.partition(d2, 1, axis=1)[:, :2]
        with np.errstate(divide="ignore", invalid="ignore"):
            g = (part[:, 1] - part[:, 0]) / part[:, 0]
        g[part[:, 0] == 0] = np.where(part[part[:, 0] == 0, 1] > 0, np.inf, 0.0)
        gap[s : s + chunk] = g
    return gap


# --------------------------------------------------------------------------- update
def sample_weights(dist: np.ndarray, total_var: float) -> np.ndarray:
    """`_calculate_exp_similarity` -- dbgsom/BaseSom.py:533-538."""
    gamma = total_var**-1
    return 1 - (1 - np.exp(-gamma * np.asarray(dist, dtype=np.float64) ** 2)) ** 0.5


def voronoi_sums(k: np.ndarray, X: np.ndarray, winners: np.ndarray, M: int):
    """Per-BMU raw sums: Sk_j = sum k_i x_i, sk_j = sum k_i, n_j = count (neuron-indexed).

    These are the quantities the device reduces (and all-reduces across GPUs); the
    reference never materialises them but its centres are Sk_j / sk_j
    (`np.average(samples[:, j], weights=weights)`, dbgsom/BaseSom.py:1052) and its
    activations are n_j (dbgsom/BaseSom.py:500-503).
    """
    X = np.asarray(X, dtype=np.float64)
    D = X.shape[1]
    Sk = np.zeros((M, D))
    if X.shape[0]:
        order = np.argsort(winners, kind="stable")
        ws = winners[order]
        starts = np.concatenate(([0], np.flatnonzero(np.diff(ws)) + 1))
        kx = (k[:, None] * X)[order]
        Sk[ws[starts]] = np.add.reduceat(kx, starts, axis=0)
    sk = np.bincount(winners, weights=k, minlength=M).astype(np.float64)
    n = np.bincount(winners, minlength=M).astype(np.float64)
    return Sk, sk, n


def voronoi_centers(k: np.ndarray, X: np.ndarray, winners: np.ndarray, M: int, pack: bool = True):
    """Step 1 of `_update_weights` + `numba_voronoi_set_centers`.

    dbgsom/BaseSom.py:488-497 and :1028-1055.  With `pack=True` (reference behaviour, quirk
    Q1) the centre of the i-th NON-EMPTY group is written to row i (`voronoi_set_centers[i,
    j]`, :1053 -- `i` is the group ordinal, not `groups[i]`), rows past the number of live
    neurons stay zero.  `pack=False` is the index-aligned variant.
    """
    Sk, sk, n = voronoi_sums(k, X, winners, M)
    live = np.flatnonzero(n > 0)
    C = np.zeros_like(Sk)
    with np.errstate(divide="ignore", invalid="ignore"):
        centres = Sk[live] / sk[live][:, None]
    if pack:
        C[: live.size] = centres
    else:
        C[live] = centres
    return C, n


def quantization_errors(winners: np.ndarray, dist: np.ndarray, M: int) -> np.ndarray:
    """`numba_quantization_error` run single-threaded == bincount.

    dbgsom/BaseSom.py:1058-1073 (the prange scatter-add is racy with >1 numba thread,
    quirk Q2; the deterministic 1-thread result equals this bincount bit for bit).
    """
    return np.bincount(winners, weights=dist, minlength=M).astype(np.float64)


# --------------------------------------------------------------------------- neighbourhood
def current_sigma(
    epoch: int,
    n_neurons: int,
    n_iter: int = 200,
    phase: str = "coarse",
    sigma_start=None,
    sigma_end=None,
    decay_function: str = "exponential",
    learning_rate: float = 0.02,
    coarse_training_frac: float = 0.5,
) -> float:
    """`_calculate_current_sigma` + decay functions -- dbgsom/BaseSom.py:863-902, :1001-1025."""
    s0 = 0.2 * sqrt(n_neurons) if sigma_start is None else sigma_start
    s1 = max(0.7, 0.05 * sqrt(n_neurons)) if sigma_end is None else sigma_end
    if phase != "coarse":
        return s1
    it = epoch / coarse_training_frac
    if decay_function == "linear":
        ratio = it / n_iter
        return s0 * (1 - ratio) + s1 * ratio
    return s1 + (s0 - s1) * exp(-learning_rate * it)


def neighborhood(hop: np.ndarray, sigma: float) -> np.ndarray:
    """`_calculate_gaussian_neighborhood` -- dbgsom/BaseSom.py:525-531."""
    return np.exp(-(np.asarray(hop, dtype=np.float64) ** 2 / (2 * sigma**2)))


def smooth(C: np.ndarray, n: np.ndarray, H: np.ndarray) -> np.ndarray:
    """Step 4 of `_update_weights` -- dbgsom/BaseSom.py:509-515.

    The reference materialises an M x M x D broadcast; this is the algebraically identical
    GEMM form W_new = ((H * n) @ C') / (H @ n), float64.
    """
    Hn = H * n[None, :]
    with np.errstate(divide="ignore", invalid="ignore"):
        return (Hn @ C) / Hn.sum(axis=1)[:, None]


def smooth_broadcast(C: np.ndarray, n: np.ndarray, H: np.ndarray) -> np.ndarray:
    """Literal shape of dbgsom/BaseSom.py:509-515 (small M only: M*M*D temporaries)."""
    inter = H[:, :, np.newaxis] * n[:, np.newaxis]
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.sum(C * inter, axis=1) / np.sum(inter, axis=1)


def epoch_step(
    X: np.ndarray,
    W: np.ndarray,
    hop: np.ndarray,
    sigma: float,
    total_var: float,
    pack: bool = True,
    bmu_fn=None,
    winners=None,
) -> dict:
    """One epoch body (dbgsom/BaseSom.py:403-407) on explicit state.

    Returns winners, dist, k, Sk, sk, n, C (centres as the reference lays them out), E,
    W_new and the convergence scalar `change` (sum_i ||W_i - W_new_i||_2, :519-520).

    `winners` (teacher forcing, tests only): take these BMU indices instead of searching, with the
    distance the reference's expansion gives for them.  The parity gate exempts samples whose two best
    float64 distances agree to 1e-6; feeding the device's choice for exactly those samples lets every
    other output of the epoch be compared without excluding anything.
    """
    W = np.asarray(W, dtype=np.float64)
    M = W.shape[0]
    fn = bmu_fn or bmu
    if winners is None:
        dist, winners = fn(X, W, 1)
    else:
        winners = np.asarray(winners, dtype=np.int64)
        dist = expansion_distance(X, W, winners)
    k = sample_weights(dist, total_var)
    Sk, sk, n = voronoi_sums(k, X, winners, M)
    C, _ = voronoi_centers(k, X, winners, M, pack=pack)
    E = quantization_errors(winners, dist, M)
    H = neighborhood(hop, sigma)
    W_new = smooth(C, n, H)
    change = float(np.sum(np.linalg.norm(W - W_new, axis=1)))
    return dict(winners=winners, dist=dist, k=k, Sk=Sk, sk=sk, n=n, C=C, E=E, W_new=W_new, change=change)


# --------------------------------------------------------------------------- helpers
def hop_matrix_grid(gx: int, gy: int) -> np.ndarray:
    """Hop counts of `nx.grid_2d_graph(gx, gy)` in node order (row-major (i, j)).

    On a full rectangular 4-connected grid the graph shortest path equals the Manhattan
    distance, which is what `nx.floyd_warshall_numpy` (dbgsom/BaseSom.py:401) returns.
    """
    ii, jj = np.meshgrid(np.arange(gx), np.arange(gy), indexing="ij")
    p = np.stack([ii.ravel(), jj.ravel()], axis=1)
    return np.abs(p[:, None, :] - p[None, :, :]).sum(axis=2).astype(np.float64)


def gmm(n: int, d: int, k: int = 64, seed: int = 0, dtype=np.float32, return_labels=False):
    """Synthetic Gaussian mixture of SURVEY.md section 8(d)."""
    rng = np.random.default_rng(seed)
    centers = rng.normal(0, 2, (k, d))
    lab = rng.integers(0, k, n)
    X = (centers[lab] + rng.normal(0, 1, (n, d))).astype(dtype)
    return (X, lab) if return_labels else X
