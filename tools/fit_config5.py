"""BASELINE.json configs[4] at reduced rows: SomVQ on embedding-scale data (D = 4096) with free growth to
~16k neurons, one GPU.

    python tools/fit_config5.py [--n 100000] [--d 4096] [--n-iter 200] [--max-neurons 16384] [--spreading-factor 0.999]

A large spreading factor makes the growing threshold small, so every boundary neuron grows in every coarse
epoch and the map reaches max_neurons well inside the coarse phase.  Prints the wall time of the whole `fit`
(upload, epochs with host growth logic, hop matrix rebuilt on the device after every growth step,
post-training passes) and a per-phase breakdown of the device time.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100_000)
    ap.add_argument("--d", type=int, default=4096)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--n-iter", type=int, default=200)
    ap.add_argument("--max-neurons", type=int, default=16384)
    ap.add_argument("--spreading-factor", type=float, default=0.999)
    ap.add_argument("--manifold", action="store_true", help="noisy 2-D sheet instead of the Gaussian mixture")
    ap.add_argument("--aligned", action="store_true", help="index-aligned centre rows instead of the reference's packing")
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    if args.manifold:
        # a noisy 2-D sheet embedded in D dimensions: a map can unfold on it, so growth really reaches max_neurons
        # (on a high-dimensional Gaussian mixture the extrapolated new prototypes win no samples and are pruned)
        basis = rng.normal(0, 1, (2, args.d)).astype(np.float32)
        X = (rng.random((args.n, 2), dtype=np.float32) * 40.0) @ basis
        X += 0.05 * rng.standard_normal((args.n, args.d), dtype=np.float32)
    else:
        centers = rng.normal(0, 2, (args.k, args.d)).astype(np.float32)
        lab = rng.integers(0, args.k, args.n)
        X = centers[lab]
        X += rng.standard_normal((args.n, args.d), dtype=np.float32)

    from dbgsom_b200 import SomVQ

    SomVQ(max_neurons=8, n_iter=3, random_state=0).fit(X[:2000])  # warm up CUDA context / library
    est = SomVQ(max_neurons=args.max_neurons, n_iter=args.n_iter, random_state=0, spreading_factor=args.spreading_factor,
                compat_pack_rows=not args.aligned, verbose=True)
    t0 = time.perf_counter()
    est.fit(X)
    dt = time.perf_counter() - t0
    epochs = est.n_iter_ + 1
    print(f"fit {dt:.2f} s, {epochs} epochs -> {epochs / dt:.2f} epochs/s, {args.n * epochs / dt:.3e} samples/s/epoch; "
          f"{len(est.neurons_)} neurons after pruning, QE {est.quantization_error_:.4f}, TE {est.topographic_error_:.4f}")
    prof = getattr(est, "fit_profile_", None)
    if prof:
        print("profile:", prof)


if __name__ == "__main__":
    main()
