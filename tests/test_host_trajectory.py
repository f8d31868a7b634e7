"""Host logic (growth rules, schedules, fitted attributes) against reference trajectories.

The estimator is driven by the test-only OracleEngine, so any difference to the
tests/golden/traj_*.npz fixtures (made by the unmodified reference) is a host-logic bug.
"""
import json
import os

import numpy as np
import pytest

import _datasets
from _oracle_engine import OracleEngine
from conftest import golden_files
from dbgsom_b200 import SomClassifier, SomVQ

TRAJ = golden_files("traj")


class Recorder(OracleEngine):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.log = dict(M=[], sigma=[], change=[], E=[], n=[])

    def epoch(self, sigma, pack_rows, entropy_error):
        self.log["M"].append(self.W.shape[0])
        self.log["sigma"].append(sigma)
        r = super().epoch(sigma, pack_rows, entropy_error)
        self.log["change"].append(r["change"])
        self.log["E"].append(np.array(r["error"], dtype=np.float64))
        self.log["n"].append(r["counts"])
        return r


def fit_with_oracle(cls, params, X, y=None):
    rec = {}

    class Est(cls):
        def _make_engine(self, distributed=None):
            e = Recorder()
            rec.setdefault("train", e)
            return e

    Est.__name__ = cls.__name__
    est = Est(**params)
    est.fit(X) if y is None else est.fit(X, y)
    return est, rec["train"]


@pytest.mark.parametrize("path", TRAJ, ids=[os.path.basename(p)[5:-4] for p in TRAJ])
def test_fit_reproduces_reference(path):
    g = np.load(path, allow_pickle=False)
    meta = json.loads(str(g["meta"]))
    X, y = _datasets.load(meta["data"])
    X = np.ascontiguousarray(X.astype(meta["cast"]))
    is_vq = meta["est"] == "vq"
    est, eng = fit_with_oracle(SomVQ if is_vq else SomClassifier, meta["params"], X, None if is_vq else y)

    # epoch-by-epoch trajectory
    np.testing.assert_array_equal(eng.log["M"], g["epoch_M"])
    np.testing.assert_allclose(eng.log["sigma"], g["epoch_sigma"], rtol=1e-14)
    np.testing.assert_allclose(np.concatenate(eng.log["n"]), g["n_flat"])
    np.testing.assert_allclose(np.concatenate(eng.log["E"]), g["E_flat"], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(eng.log["change"], g["epoch_change"], rtol=1e-6, atol=1e-9)

    # fitted attributes
    assert est.n_iter_ == int(g["n_iter_"])
    assert bool(est.converged_) == bool(g["converged"])
    np.testing.assert_array_equal(np.array(est.neurons_), g["neurons"])
    np.testing.assert_allclose(est.weights_, g["weights"], rtol=1e-8, atol=1e-9)
    np.testing.assert_array_equal(est._distance_matrix, g["distance_matrix"])
    assert est.quantization_error_ == pytest.approx(float(g["quantization_error"]), rel=1e-9)
    assert est.topographic_error_ == pytest.approx(float(g["topographic_error"]), rel=1e-12)
    assert est.growing_threshold_ == pytest.approx(float(g["growing_threshold"]), rel=1e-12)
    assert est.n_features_in_ == X.shape[1]
    edges = np.array(sorted(tuple(sorted(e)) for e in est.som_.edges), dtype=np.int64).reshape(-1, 2, 2)
    np.testing.assert_array_equal(edges, g["edges"])
    np.testing.assert_array_equal(est._extract_values_from_graph("label"), g["node_label"])
    np.testing.assert_allclose(est._extract_values_from_graph("error"), g["node_error"], rtol=1e-8, atol=1e-10)
    alive = g["hit_count_pre"] > 0
    np.testing.assert_array_equal(est._extract_values_from_graph("hit_count"), g["hit_count_pre"][alive])
    np.testing.assert_allclose(est._extract_values_from_graph("density"), g["density_pre"][alive], rtol=1e-8)
    np.testing.assert_allclose(
        est._extract_values_from_graph("average_distance"), g["avgdist_pre"][alive], rtol=1e-9
    )
    np.testing.assert_array_equal(
        est._extract_values_from_graph("epoch_created"), g["epoch_created_pre"][alive]
    )
    np.testing.assert_allclose(est.transform(X[:20]), g["transform_head"], rtol=1e-6, atol=1e-9)
    if is_vq:
        np.testing.assert_array_equal(est.labels_, g["labels"])
        np.testing.assert_array_equal(est.predict(X[:200]), g["predict_head"])
    else:
        np.testing.assert_array_equal(est.classes_, g["classes"])
        np.testing.assert_allclose(est._extract_values_from_graph("probabilities"), g["probabilities"])
        np.testing.assert_array_equal(est.predict(X[:200]), g["predict_head"])
        np.testing.assert_allclose(est.predict_proba(X[:50]), g["proba_head"], rtol=1e-6, atol=1e-9)
        assert est.score(X, y) == pytest.approx(float(g["score"]))


FITSTEP = golden_files("fitstep")


@pytest.mark.parametrize("path", FITSTEP, ids=[os.path.basename(p)[8:-4] for p in FITSTEP])
def test_config2_fit_reproduces_reference(path):
    """BASELINE.json configs[1] (SomClassifier, 70000 x 784, ten classes, reduced n_iter): the host logic driven by
    the oracle engine grows the reference's map epoch by epoch and ends with its neurons, prototypes and score."""
    g = np.load(path, allow_pickle=False)
    meta = json.loads(str(g["meta"]))
    X, y = _datasets.load(meta["data"])
    X = np.ascontiguousarray(X.astype(meta["cast"]))
    est, eng = fit_with_oracle(SomClassifier, meta["params"], X, y)
    np.testing.assert_array_equal(eng.log["M"], g["epoch_M"])
    np.testing.assert_allclose(np.concatenate(eng.log["n"]), g["n_flat"])
    np.testing.assert_allclose(np.concatenate(eng.log["E"]), g["E_flat"], rtol=1e-6, atol=1e-8)
    np.testing.assert_array_equal(np.array(est.neurons_), g["neurons"])
    scale = np.abs(g["weights"]).max()
    assert np.abs(est.weights_ - g["weights"]).max() / scale < 1e-6
    assert est.n_iter_ == int(g["n_iter_"])
    assert est.quantization_error_ == pytest.approx(float(g["quantization_error"]), rel=1e-7)
    assert est.topographic_error_ == pytest.approx(float(g["topographic_error"]), abs=1e-12)
    np.testing.assert_array_equal(est._extract_values_from_graph("label"), g["node_label"])
    np.testing.assert_array_equal(est.classes_, g["classes"])
