"""BASELINE.json configs[1]: SomClassifier on Fashion-MNIST-shaped synthetic 70000 x 784, growth to ~400 neurons.

    python tools/fit_config2.py [--n-iter 240] [--max-neurons 400]

Prints wall time of the whole `fit` (upload, all epochs with host growth logic, post-training passes),
epochs/s, samples/s/epoch and the fitted map's size / metrics.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=70000)
    ap.add_argument("--d", type=int, default=784)
    ap.add_argument("--n-iter", type=int, default=240)
    ap.add_argument("--max-neurons", type=int, default=400)
    ap.add_argument("--aligned", action="store_true", help="index-aligned centre rows instead of the reference's packing")
    ap.add_argument("--backend", default="auto")
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    centers = rng.normal(0, 2, (10, args.d))
    y = rng.integers(0, 10, args.n)
    X = (centers[y] + rng.normal(0, 1, (args.n, args.d))).astype(np.float32)

    from dbgsom_b200 import SomClassifier

    est = SomClassifier(max_neurons=args.max_neurons, n_iter=args.n_iter, random_state=0,
                        compat_pack_rows=not args.aligned, bmu_backend=args.backend)
    SomClassifier(max_neurons=8, n_iter=3, random_state=0).fit(X[:2000], y[:2000])  # warm up CUDA context / library
    t0 = time.perf_counter()
    est.fit(X, y)
    dt = time.perf_counter() - t0
    epochs = est.n_iter_ + 1
    t1 = time.perf_counter()
    acc = est.score(X[:5000], y[:5000])
    t_pred = time.perf_counter() - t1
    print(f"fit {dt:.2f} s, {epochs} epochs -> {epochs / dt:.1f} epochs/s, {args.n * epochs / dt:.3e} samples/s/epoch; "
          f"{len(est.neurons_)} neurons after pruning, QE {est.quantization_error_:.4f}, TE {est.topographic_error_:.4f}, "
          f"accuracy(5000) {acc:.3f} (predict {t_pred:.2f} s, host LARS)")


if __name__ == "__main__":
    main()
