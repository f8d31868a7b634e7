"""The LARS-lasso restatement (dbgsom_b200/csrc/lars_core.cuh, run per CUDA thread by dbgsom_sparse_code) against
scikit-learn's own solver, on the CPU through a test-only host build of the same header."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
from sklearn.decomposition import SparseCoder
from sklearn.preprocessing import normalize

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lars_host(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("lars") / "lars_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "lars_host_shim.cpp")])
    lib = ctypes.CDLL(out)
    lib.lars_host.restype = ctypes.c_int

    def run(Wn, Xn, A=None, max_iter=1000):
        m, d = Wn.shape
        gram = np.ascontiguousarray(Wn @ Wn.T)
        cov = np.ascontiguousarray(Xn @ Wn.T)
        n = Xn.shape[0]
        code = np.zeros((n, m))
        status = np.zeros(n, dtype=np.int32)
        lib.lars_host(gram.ctypes.data_as(ctypes.c_void_p), m, cov.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(n), d,
                      max_iter, A or m, code.ctypes.data_as(ctypes.c_void_p), status.ctypes.data_as(ctypes.c_void_p))
        return code, status

    return run


def sklearn_code(Wn, Xn):
    coder = SparseCoder(dictionary=Wn, positive_code=True, transform_alpha=0, transform_algorithm="lasso_lars")
    return coder.transform(Xn)


def sheet(m_side, d, rng, smooth=2.0):
    """A smooth two-dimensional sheet of prototypes in d dimensions (what a trained map looks like)."""
    u, v = np.meshgrid(np.linspace(0, 1, m_side), np.linspace(0, 1, m_side), indexing="ij")
    basis = rng.normal(size=(6, d))
    feats = np.stack([u, v, np.sin(smooth * u), np.cos(smooth * v), u * v, np.ones_like(u)], axis=-1).reshape(-1, 6)
    return feats @ basis + 0.05 * rng.normal(size=(m_side * m_side, d))


@pytest.mark.parametrize("case", ["random_tall", "random_wide", "sheet", "clustered", "tiny"])
def test_matches_sklearn_lasso_lars(lars_host, case):
    rng = np.random.default_rng(hash(case) % 1000)
    if case == "random_tall":
        W, X = rng.normal(size=(30, 64)), rng.normal(size=(200, 64))
    elif case == "random_wide":
        W, X = rng.normal(size=(80, 16)) + 1.0, rng.normal(size=(150, 16)) + 1.0
    elif case == "sheet":
        W = sheet(7, 48, rng)
        X = W[rng.integers(0, 49, 300)] + 0.3 * rng.normal(size=(300, 48))
    elif case == "clustered":
        centers = rng.normal(0, 2, (6, 32))
        W = centers[rng.integers(0, 6, 40)] + 0.2 * rng.normal(size=(40, 32))
        X = centers[rng.integers(0, 6, 250)] + rng.normal(size=(250, 32))
    else:
        W, X = rng.normal(size=(4, 5)), rng.normal(size=(50, 5))
    Wn, Xn = normalize(W), normalize(X)
    ref = sklearn_code(Wn, Xn)
    code, status = lars_host(Wn, Xn)
    assert not (status & 1).any()
    np.testing.assert_allclose(code, ref, rtol=1e-7, atol=1e-10)
    assert (code >= -1e-12).all()  # a coefficient leaving the active set lands on zero up to rounding, as in sklearn
    assert ((code > 1e-12) == (ref > 1e-12)).mean() > 0.999


def test_capacity_flag_and_rerun(lars_host):
    rng = np.random.default_rng(5)
    W, X = rng.normal(size=(40, 24)), rng.normal(size=(60, 24))
    Wn, Xn = normalize(W), normalize(X)
    ref = sklearn_code(Wn, Xn)
    code, status = lars_host(Wn, Xn, A=3)
    over = (status & 1) != 0
    assert over.any() and not over.all() or over.all()
    fine = ~over
    np.testing.assert_allclose(code[fine], ref[fine], rtol=1e-7, atol=1e-10)
    code2, status2 = lars_host(Wn, Xn, A=40)
    assert not (status2 & 1).any()
    np.testing.assert_allclose(code2, ref, rtol=1e-7, atol=1e-10)


def test_matches_reference_transform_fixtures(lars_host):
    """`transform(X[:20])` of the unmodified reference (tests/golden/traj_*.npz) from its fitted prototypes."""
    import json

    import _datasets
    from conftest import golden_files

    for path in golden_files("traj"):
        g = np.load(path, allow_pickle=False)
        meta = json.loads(str(g["meta"]))
        X, _ = _datasets.load(meta["data"])
        X = np.ascontiguousarray(X.astype(meta["cast"]))[:20].astype(np.float64)
        code, status = lars_host(normalize(g["weights"]), normalize(X))
        assert not (status & 1).any()
        np.testing.assert_allclose(code, g["transform_head"], rtol=1e-6, atol=1e-9, err_msg=path)
