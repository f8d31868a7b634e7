"""BASELINE.json configs[4]: SomVQ on embedding-scale data (D = 4096) with free growth to ~16k neurons, on one GPU
at reduced rows or sharded over the GPUs of a box.

    python tools/fit_config5.py [--n 100000] [--d 4096] [--n-iter 200] [--max-neurons 16384] [--spreading-factor 0.999]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29741 \
        tools/fit_config5.py --distributed --rows 625000 --manifold --aligned     # rows PER RANK (torchrun claims --n)

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
    ap.add_argument("--n", "--rows", dest="n", type=int, default=100_000, help="rows (per rank with --distributed)")
    ap.add_argument("--d", type=int, default=4096)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--n-iter", type=int, default=200)
    ap.add_argument("--max-neurons", type=int, default=16384)
    ap.add_argument("--spreading-factor", type=float, default=0.999)
    ap.add_argument("--coarse-frac", type=float, default=0.5, help="share of the epochs in the coarse (growing) phase")
    ap.add_argument("--manifold", action="store_true", help="noisy 2-D sheet instead of the Gaussian mixture")
    ap.add_argument("--aligned", action="store_true", help="index-aligned centre rows instead of the reference's packing")
    ap.add_argument("--distributed", action="store_true", help="one rank per GPU under torchrun; --n rows per rank")
    ap.add_argument("--json", default=None, help="rank 0 writes a one-line JSON record here")
    args = ap.parse_args()
    rank, world, device = 0, 1, "cuda"
    if args.distributed:
        import torch
        import torch.distributed as dist

        local = int(os.environ["LOCAL_RANK"])
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rank, world, device = dist.get_rank(), dist.get_world_size(), f"cuda:{local}"
    rng = np.random.default_rng(0)          # the geometry (basis / centres) is the same on every rank ...
    rng_rows = np.random.default_rng(1000 + rank)  # ... the rows of a shard are its own
    if args.manifold:
        # a noisy 2-D sheet embedded in D dimensions: a map can unfold on it, so growth really reaches max_neurons
        # (on a high-dimensional Gaussian mixture the extrapolated new prototypes win no samples and are pruned)
        basis = rng.normal(0, 1, (2, args.d)).astype(np.float32)
        X = (rng_rows.random((args.n, 2), dtype=np.float32) * 40.0) @ basis
        for s in range(0, args.n, 1 << 16):  # noise in row chunks: no second matrix of the shard's size on the host
            e = min(args.n, s + (1 << 16))
            X[s:e] += 0.05 * rng_rows.standard_normal((e - s, args.d), dtype=np.float32)
    else:
        centers = rng.normal(0, 2, (args.k, args.d)).astype(np.float32)
        lab = rng_rows.integers(0, args.k, args.n)
        X = centers[lab]
        for s in range(0, args.n, 1 << 16):
            e = min(args.n, s + (1 << 16))
            X[s:e] += rng_rows.standard_normal((e - s, args.d), dtype=np.float32)

    from dbgsom_b200 import SomVQ

    SomVQ(max_neurons=8, n_iter=3, random_state=0, device=device).fit(X[:2000])  # warm up CUDA context / library
    est = SomVQ(max_neurons=args.max_neurons, n_iter=args.n_iter, random_state=0, spreading_factor=args.spreading_factor,
                compat_pack_rows=not args.aligned, verbose=rank == 0, device=device, distributed=args.distributed,
                coarse_training_frac=args.coarse_frac)
    if args.distributed:
        dist.barrier()
    t0 = time.perf_counter()
    est.fit(X)
    if args.distributed:
        dist.barrier()
    dt = time.perf_counter() - t0
    epochs = est.n_iter_ + 1
    rows_total = args.n * world
    if rank == 0:
        grown = int(getattr(est, "_max_map_size", 0)) or None
        print(f"fit {dt:.2f} s on {world} GPU(s), {rows_total} x {args.d} rows, {epochs} epochs -> {epochs / dt:.2f} epochs/s, "
              f"{rows_total * epochs / dt:.3e} samples/s/epoch; map grew to {grown} neurons, {len(est.neurons_)} after pruning, "
              f"QE {est.quantization_error_:.4f}, TE {est.topographic_error_:.4f}")
        prof = getattr(est, "fit_profile_", None)
        if prof:
            print("profile:", prof)
        if args.json:
            import json

            rec = {"tool": "fit_config5", "gpus": world, "rows_total": rows_total, "d": args.d, "epochs": int(epochs), "coarse_training_frac": args.coarse_frac,
                   "fit_s": dt, "epochs_per_s": epochs / dt, "samples_per_s_per_epoch": rows_total * epochs / dt,
                   "neurons_grown_to": grown, "neurons_after_pruning": len(est.neurons_), "max_neurons": args.max_neurons,
                   "data": "noisy 2-D sheet in D dimensions" if args.manifold else "Gaussian mixture",
                   "aligned_centres": bool(args.aligned), "quantization_error": float(est.quantization_error_),
                   "topographic_error": float(est.topographic_error_), "profile": prof}
            with open(args.json, "w") as f:
                f.write(json.dumps(rec) + "\n")
    if args.distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
