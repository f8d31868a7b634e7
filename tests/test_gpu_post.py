"""Post-training passes and the hop matrix on the device against
numpy / scipy restatements of the reference lines they replace.  All tests need a B200 (`-m gpu`)."""
import numpy as np
import pytest

import _datasets
from dbgsom_b200.topology import MapTopology
from oracle import som_oracle as O

pytestmark = pytest.mark.gpu


def engine(**kw):
    from dbgsom_b200.engine import DeviceEngine

    return DeviceEngine(**kw)


def grown_topology(steps=40, seed=0):
    """An irregular map: random errors drive the reference growth rules for a number of passes."""
    rng = np.random.default_rng(seed)
    t = MapTopology.initial_square()
    for e in range(steps):
        t.error[:] = rng.random(len(t)) * 10
        t.distribute_errors(6.0)
        t.grow(6.0, e)
    return t


@pytest.mark.parametrize("steps", [0, 5, 40, 120])
def test_hops_match_host_bfs(steps):
    topo = grown_topology(steps)
    e = engine()
    e.set_hops_from_topology(topo)
    ref = topo.hop_matrix_u16()
    np.testing.assert_array_equal(e.hops_host(), ref)
    assert e.hop_max == int(ref[ref != 0xFFFF].max())
    e.close()


def test_hops_disconnected_and_full_grid():
    topo = MapTopology.full_grid(64, 64)
    e = engine()
    e.set_hops_from_topology(topo)
    p = topo.positions()
    np.testing.assert_array_equal(e.hops_host(), np.abs(p[:, None, :] - p[None, :, :]).sum(axis=2).astype(np.uint16))
    # removing a column of a 5 x 9 grid splits it: unreachable pairs are 0xFFFF (dbgsom/BaseSom.py:235)
    g = MapTopology.full_grid(5, 9)
    cut = g.without(np.array([i for i, q in enumerate(g.pos) if q[1] == 4]))
    e.set_hops_from_topology(cut)
    ref = cut.hop_matrix_u16()
    assert (ref == 0xFFFF).any()
    np.testing.assert_array_equal(e.hops_host(), ref)
    e.close()


@pytest.mark.parametrize("shape", [(30000, 64, 36), (200000, 32, 300), (5000, 200, 9)])
def test_final_statistics_match_numpy(shape):
    from scipy.spatial.distance import cdist

    n, d, side2 = shape
    X = _datasets.gmm(n, d, 8, 5)
    gx = int(np.sqrt(side2))
    topo = MapTopology.full_grid(gx, side2 // gx)
    m = len(topo)
    W0 = X[np.random.default_rng(1).choice(n, m, replace=False)].astype(np.float64)
    e = engine(bmu_backend="simt")
    e.load_data(X, None, 0)
    e.set_map(W0)
    e.set_hops_from_topology(topo)
    e.epoch(1.3, True, False)            # W_prev = W0, W = updated prototypes
    W1 = e.weights()
    st = e.final_statistics(topo.positions(), topo.degrees())
    X64 = X.astype(np.float64)
    dist2, idx2 = O.bmu(X64, W0, 2)
    gap = O.relative_gap(X64, W0)
    deg = topo.degrees()
    avg = (cdist(W1, W1) * deg[None, :]).sum(axis=1) / deg.sum()
    np.testing.assert_allclose(st["avg_dist"], avg, rtol=1e-12)
    bw = avg.mean()
    assert st["n_rows"] == m
    # Never skipped: samples whose two best float64 distances agree to 1e-6 (the parity gate's exempt set) may swap
    # their first and second BMU on the device; every per-neuron statistic is bracketed by the two assignments of
    # those samples and must be exact for all the others.
    near = gap < 1e-6
    assert near.mean() < 0.01
    pos = topo.positions().astype(np.float64)
    sep = pos[idx2[:, 0]] - pos[idx2[:, 1]]
    # second BMUs may legitimately differ inside near-ties of the 2nd/3rd distance: compare loosely there
    te = np.count_nonzero(np.sqrt((sep * sep).sum(axis=1)) > 1.5)
    assert abs(st["te_count"] - te) <= max(2, 1e-4 * n) + near.sum()
    first, second = idx2[:, 0], idx2[:, 1]
    lo = np.bincount(first[~near], minlength=m)
    hi = lo + np.bincount(first[near], minlength=m) + np.bincount(second[near], minlength=m)
    assert (st["hits"] >= lo).all() and (st["hits"] <= hi).all() and st["hits"].sum() == n
    kern = np.exp(-(dist2[:, 0] ** 2) / (2 * bw**2)) / (bw * np.sqrt(2 * np.pi))
    dlo = np.bincount(first[~near], weights=kern[~near], minlength=m)
    dhi = dlo + np.bincount(first[near], weights=kern[near], minlength=m) + np.bincount(second[near], weights=kern[near], minlength=m)
    assert (st["dens_sum"] >= dlo * (1 - 1e-9) - 1e-300).all() and (st["dens_sum"] <= dhi * (1 + 1e-6) + 1e-300).all()
    if not near.any():
        np.testing.assert_array_equal(st["hits"], np.bincount(first, minlength=m))
        np.testing.assert_allclose(st["dens_sum"], np.bincount(first, weights=kern, minlength=m), rtol=1e-9, atol=1e-300)
    assert st["qe_sum"] == pytest.approx(dist2[:, 0].sum(), rel=1e-9)
    np.testing.assert_allclose(st["weights"], W1)
    e.close()


def test_label_histogram_matches_numpy():
    n, d, c = 60000, 48, 7
    X = _datasets.gmm(n, d, c, 2)
    y = np.random.default_rng(3).integers(0, c, n)
    topo = MapTopology.full_grid(5, 6)
    m = len(topo)
    W0 = X[np.random.default_rng(4).choice(n, m, replace=False)].astype(np.float64)
    e = engine(bmu_backend="simt")
    e.load_data(X, y, c)
    e.set_map(W0)
    e.final_winners()
    win = e.winners_host()
    _, ref = O.bmu(X.astype(np.float64), W0, 1)
    gap = O.relative_gap(X.astype(np.float64), W0)
    np.testing.assert_array_equal(win[gap >= 1e-6], ref.reshape(-1)[gap >= 1e-6])
    counts, first = e.label_histogram(c)
    flat = win * c + y
    np.testing.assert_array_equal(counts.reshape(-1), np.bincount(flat, minlength=m * c))
    exp_first = np.full(m * c, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(exp_first, flat, np.arange(n, dtype=np.int64))
    np.testing.assert_array_equal(first.reshape(-1), exp_first)
    e.close()


@pytest.mark.parametrize("case", ["sheet", "wide", "dense_paths"])
def test_sparse_code_matches_sklearn(case):
    """BaseSom.transform on the device (dbgsom_sparse_code) against the reference's scikit-learn call."""
    from sklearn.decomposition import SparseCoder
    from sklearn.preprocessing import normalize

    rng = np.random.default_rng(11)
    if case == "sheet":   # smooth trained-map-like dictionary: 2-7 active atoms per sample
        u, v = np.meshgrid(np.linspace(0, 1, 9), np.linspace(0, 1, 9), indexing="ij")
        feats = np.stack([u, v, np.sin(2 * u), np.cos(2 * v), u * v, np.ones_like(u)], axis=-1).reshape(-1, 6)
        W = feats @ rng.normal(size=(6, 64)) + 0.05 * rng.normal(size=(81, 64))
        X = W[rng.integers(0, 81, 6000)] + 0.3 * rng.normal(size=(6000, 64))
    elif case == "wide":  # more atoms than features
        W, X = rng.normal(size=(120, 24)) + 1.0, rng.normal(size=(3000, 24)) + 1.0
    else:                 # long paths: active sets beyond the first Cholesky capacity (32) -> rerun at 128
        W, X = rng.normal(size=(100, 96)), rng.normal(size=(700, 96))
    Wn, Xn = normalize(W), normalize(X)
    ref = SparseCoder(dictionary=Wn, positive_code=True, transform_alpha=0, transform_algorithm="lasso_lars").transform(Xn)
    e = engine()
    code = e.sparse_code(Xn, Wn)
    e.close()
    if case == "dense_paths":
        assert (np.count_nonzero(ref > 1e-12, axis=1) > 32).any()
    np.testing.assert_allclose(code, ref, rtol=1e-6, atol=1e-9)


def test_sparse_code_float32_samples():
    """float32 samples: scikit-learn then runs the Cholesky part of the path in float32 (its Gram matrix is cast to
    the dtype of X, sklearn/decomposition/_dict_learning.py `_sparse_encode`), the device path stays in float64.
    On a well-conditioned dictionary both agree to float32 round-off."""
    from sklearn.decomposition import SparseCoder
    from sklearn.preprocessing import normalize

    rng = np.random.default_rng(12)
    W, X = rng.normal(size=(30, 64)), rng.normal(size=(2000, 64)).astype(np.float32)
    Wn, Xn = normalize(W), normalize(X)
    ref = SparseCoder(dictionary=Wn, positive_code=True, transform_alpha=0, transform_algorithm="lasso_lars").transform(Xn)
    e = engine()
    code = e.sparse_code(Xn, Wn)
    e.close()
    np.testing.assert_allclose(code, ref, rtol=0, atol=2e-6)
