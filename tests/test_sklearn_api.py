"""scikit-learn contract of the estimators (CPU: driven by the test-only oracle engine)."""
import numpy as np
import pytest
from sklearn.base import clone
from sklearn.pipeline import make_pipeline
from sklearn.preprocessing import StandardScaler

import _datasets
from _oracle_engine import OracleEngine
from dbgsom_b200 import SomClassifier, SomVQ
from dbgsom_b200.BaseSom import BaseSom, sigma_exponential, sigma_linear

REFERENCE_DEFAULTS = dict(
    n_iter=200, convergence_iter=1, spreading_factor=0.5, sigma_start=None, sigma_end=None, vertical_growth=False,
    decay_function="exponential", learning_rate=0.02, verbose=False, coarse_training_frac=0.5, random_state=None,
    convergence_treshold=10**-5, max_neurons=100, metric="euclidean", threshold_method="se",
    growth_criterion="quantization_error", min_samples_vertical_growth=100, n_jobs=1,
)


def with_oracle(cls):
    return type(cls.__name__, (cls,), {"_make_engine": lambda self, distributed=None: OracleEngine()})


def test_hyperparameters_match_the_reference():
    """Same names and defaults as dbgsom/BaseSom.py:42-80 (plus the device-side additions)."""
    for cls in (SomVQ, SomClassifier):
        params = cls().get_params()
        for k, v in REFERENCE_DEFAULTS.items():
            assert params[k] == v, k
        assert set(params) - set(REFERENCE_DEFAULTS) == {"device", "compat_pack_rows", "bmu_backend", "distributed", "bound_scale", "strict_ties"}


def test_clone_and_set_params_roundtrip():
    est = SomVQ(max_neurons=17, sigma_end=0.3, random_state=5, bmu_backend="simt")
    twin = clone(est)
    assert twin.get_params() == est.get_params() and twin is not est
    twin.set_params(n_iter=7)
    assert twin.n_iter == 7 and est.n_iter == 200


def test_argument_validation():
    X = np.random.default_rng(0).normal(size=(50, 4))
    with pytest.raises(ValueError, match="Decay function"):
        with_oracle(SomVQ)(decay_function="cosine").fit(X)
    with pytest.raises(ValueError, match="threshold_method"):
        with_oracle(SomVQ)(threshold_method="x").fit(X)
    with pytest.raises(ValueError, match="growth_criterion"):
        with_oracle(SomVQ)(growth_criterion="x").fit(X)
    with pytest.raises(NotImplementedError):
        with_oracle(SomVQ)(vertical_growth=True).fit(X)
    with pytest.raises(ValueError):  # sklearn input validation, like the reference (ensure_min_samples=4)
        with_oracle(SomVQ)().fit(X[:3])
    with pytest.raises(ValueError):
        with_oracle(SomVQ)().fit(np.array([[np.nan, 1.0]] * 10))


def test_fit_predict_and_pipeline():
    X, y = _datasets.load("rings3d")
    vq = with_oracle(SomVQ)(n_iter=20, max_neurons=20, random_state=0)
    labels = vq.fit_predict(X)
    np.testing.assert_array_equal(labels, vq.labels_)
    assert labels.max() < len(vq.neurons_) and vq.weights_.shape == (len(vq.neurons_), 3)
    codes = vq.transform(X[:10])
    assert codes.shape == (10, len(vq.neurons_)) and (codes >= 0).all()
    pipe = make_pipeline(StandardScaler(), with_oracle(SomClassifier)(n_iter=20, max_neurons=20, random_state=0))
    pipe.fit(X, y)
    assert pipe.score(X, y) > 0.8
    proba = pipe.predict_proba(X[:5])
    np.testing.assert_allclose(proba.sum(axis=1), 1.0)
    assert set(pipe.predict(X[:50])) <= set(np.unique(y))


def test_sigma_schedules_follow_the_reference_formulas():
    assert sigma_linear(4.0, 1.0, 100, 50) == pytest.approx(2.5)
    assert sigma_exponential(4.0, 1.0, 100, 10, 0.02) == pytest.approx(1.0 + 3.0 * np.exp(-0.2))
    # the value the reference's own (failing) unit test documents for its exponential_decay:
    # dbgsom/test_dbgsom_.py:13-29 expects 0.125, the function returns 0.14097959895689505
    assert sigma_exponential(1.0, 0.0, 100, 0.5 * 100 * 0 + 0, 0.02) == 1.0


def test_nan_prototypes_raise_like_the_reference():
    class Broken(OracleEngine):
        def epoch(self, *a, **k):
            r = super().epoch(*a, **k)
            r["change"] = float("nan")
            return r

    Est = type("SomVQ", (SomVQ,), {"_make_engine": lambda self, distributed=None: Broken()})
    with pytest.raises(ValueError, match="NaN"):
        Est(n_iter=5).fit(np.random.default_rng(0).normal(size=(40, 3)))


def test_product_engine_is_the_default_and_needs_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        SomVQ(n_iter=2).fit(np.random.default_rng(0).normal(size=(40, 3)))
    assert BaseSom._make_engine.__module__ == "dbgsom_b200.BaseSom"


def test_entropy_growth_without_labels_is_an_error():
    """SomVQ has no y: the reference dies in its first epoch (y[winners == j] with y = None); say why."""
    X = np.random.default_rng(0).normal(size=(50, 5))
    with pytest.raises(ValueError, match="entropy"):
        with_oracle(SomVQ)(growth_criterion="entropy").fit(X)


def test_survivor_without_samples_on_the_final_map_gets_zero_probabilities():
    """dbgsom/SomClassifier.py:136-152: a neuron that survives dead-neuron removal (hit_count > 0 on the pre-update
    prototypes) but wins nothing on the updated map gets label -1 and an ALL-ZERO probability row."""
    import networkx as nx

    class Hist:
        def label_histogram(self, c):
            counts = np.array([[3.0, 1.0, 0.0], [0.0, 0.0, 0.0]])
            first = np.array([[0, 2, 9], [9, 9, 9]], dtype=np.int64)
            return counts, first

    est = SomClassifier()
    est.classes_ = np.array([0, 1, 2])
    est.neurons_ = [(0, 0), (0, 1)]
    est.som_ = nx.Graph()
    est.som_.add_node((0, 0), hit_count=4.0)
    est.som_.add_node((0, 1), hit_count=2.0)
    est._label_prototypes(None, Hist())
    assert est.som_.nodes[(0, 0)]["label"] == 0
    np.testing.assert_allclose(est.som_.nodes[(0, 0)]["probabilities"], [0.75, 0.25, 0.0])
    assert est.som_.nodes[(0, 1)]["label"] == -1
    np.testing.assert_array_equal(est.som_.nodes[(0, 1)]["probabilities"], [0.0, 0.0, 0.0])
