"""N > 1 host path on CPU: two gloo ranks, each with half of the samples, must build the same map as
one process with all samples (only the per-neuron partial sums cross ranks)."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, out_dir):
    sys.path.insert(0, os.path.dirname(HERE))
    sys.path.insert(0, HERE)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist

    import _datasets
    from _oracle_engine import ShardedOracleEngine
    from dbgsom_b200 import SomClassifier, SomVQ
    from dbgsom_b200.engine import Comm

    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = Comm(True)
        X, y = _datasets.load("digits")
        n = X.shape[0]
        lo, hi = rank * n // world, (rank + 1) * n // world

        def factory(self, distributed=None):
            return ShardedOracleEngine(comm) if distributed is not False else __import__("_oracle_engine").OracleEngine()

        cls = SomVQ if kind == "vq" else SomClassifier
        Est = type(cls.__name__, (cls,), {"_make_engine": factory})
        if kind == "clf_sorted_unseeded":
            # shards with DIFFERENT class sets (samples sorted by label) and no seed: the start rows must come from
            # one rank and classes_ from the union of the shards, or the ranks train different maps
            order = np.argsort(y, kind="stable")
            X, y = X[order], y[order]
            est = Est(random_state=None, n_iter=12, distributed=True)
            est.fit(X[lo:hi], y[lo:hi])
            assert set(np.unique(y[lo:hi])) != set(np.unique(y)), "the shard should miss classes"
            np.savez(
                os.path.join(out_dir, f"rank{rank}.npz"), neurons=np.array(est.neurons_), weights=est.weights_,
                classes=est.classes_, labels=est._extract_values_from_graph("label"),
                prob=est._extract_values_from_graph("probabilities"), qe=est.quantization_error_,
            )
            return
        est = Est(random_state=0, n_iter=30, distributed=True)
        est.fit(X[lo:hi]) if kind == "vq" else est.fit(X[lo:hi], y[lo:hi])
        np.savez(
            os.path.join(out_dir, f"rank{rank}.npz"), neurons=np.array(est.neurons_), weights=est.weights_,
            qe=est.quantization_error_, te=est.topographic_error_, hits=est._extract_values_from_graph("hit_count"),
            labels=est._extract_values_from_graph("label"),
            local=est.labels_ if kind == "vq" else est.predict(X[lo : lo + 50]),
        )
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["vq", "clf"])
def test_two_ranks_equal_one_process(tmp_path, kind):
    import torch.multiprocessing as mp

    import _datasets
    from _oracle_engine import OracleEngine
    from dbgsom_b200 import SomClassifier, SomVQ

    mp.spawn(_worker, args=(2, _free_port(), kind, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    # both ranks hold the same replicated map
    for key in ("neurons", "weights", "qe", "te", "hits", "labels"):
        np.testing.assert_array_equal(r0[key], r1[key])

    X, y = _datasets.load("digits")
    cls = SomVQ if kind == "vq" else SomClassifier
    Est = type(cls.__name__, (cls,), {"_make_engine": lambda self, distributed=None: OracleEngine()})
    one = Est(random_state=0, n_iter=30)
    one.fit(X) if kind == "vq" else one.fit(X, y)
    np.testing.assert_array_equal(r0["neurons"], np.array(one.neurons_))
    np.testing.assert_allclose(r0["weights"], one.weights_, rtol=1e-9, atol=1e-9)
    assert float(r0["qe"]) == pytest.approx(one.quantization_error_, rel=1e-9)
    assert float(r0["te"]) == pytest.approx(one.topographic_error_, rel=1e-12)
    np.testing.assert_array_equal(r0["hits"], one._extract_values_from_graph("hit_count"))
    np.testing.assert_array_equal(r0["labels"], one._extract_values_from_graph("label"))
    if kind == "vq":
        n = X.shape[0]
        np.testing.assert_array_equal(np.concatenate([r0["local"], r1["local"]]), one.labels_)


def test_two_ranks_unseeded_with_label_sorted_shards(tmp_path):
    """with random_state=None every rank drew its own start rows, and classes_ came from the
    local shard.  Both ranks must end with the same map, the union of the classes and a finite fit."""
    import torch.multiprocessing as mp

    import _datasets

    mp.spawn(_worker, args=(2, _free_port(), "clf_sorted_unseeded", str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "rank0.npz"), np.load(tmp_path / "rank1.npz")
    for key in ("neurons", "weights", "classes", "labels", "prob", "qe"):
        np.testing.assert_array_equal(r0[key], r1[key])
    _, y = _datasets.load("digits")
    np.testing.assert_array_equal(r0["classes"], np.unique(y))
    assert r0["prob"].shape[1] == 10 and np.isfinite(r0["weights"]).all()
    # the four start prototypes were real samples (a sum of unrelated rows would sit far outside the data)
    assert r0["weights"].min() >= -1e-9 and r0["weights"].max() <= 16 + 1e-9
