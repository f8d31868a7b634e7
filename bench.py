#!/usr/bin/env python
"""Benchmark of the batch-SOM training epoch (BASELINE.json metric: training samples/s per epoch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c4|c2|small] [--impl reference]

One "step" = one full training epoch (BMU search -> per-BMU accumulation -> [all-reduce] ->
neighbourhood smoothing -> read-back of per-neuron error/count/change) of a fixed side x side map on
synthetic Gaussian-mixture data, prototypes evolving from epoch to epoch exactly as in `fit`
(reference semantics incl. packed centre rows, sigma schedule of a 200-epoch fit).

Workloads; `c3` is the default at every GPU count (weak scaling: each GPU holds
10M x 256 samples, the only collective is the all-reduce of the M x D partial sums):
    c3     10M x 256 per GPU, 64 x 64 map           (BASELINE.json configs[2], the roofline config)
    c4     100M x 128 in total, sharded, 64 x 64    (configs[3], strong scaling)
    c2     70000 x 784, 20 x 20 map                 (configs[1]; latency-bound)
    c5     5M x 4096 in total, sharded, 128 x 128   (configs[4] at its final map size; needs 8 GPUs or --rows)
    small  200k x 64, 16 x 16 map                   (smoke-sized)

`--impl reference` times the CPU restatement of the reference's epoch (oracle/, numpy + the same
scikit-learn call the reference makes) on the host cores, on a bounded row sample of the workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    #          rows per GPU (None = total/gpus), total rows, D, side, mixture components
    "c3": dict(per_gpu=10_000_000, total=None, d=256, side=64, k=64, name="SomVQ epoch, GMM 10M x 256 per GPU, fixed 64x64 map"),
    "c4": dict(per_gpu=None, total=100_000_000, d=128, side=64, k=64, name="SomVQ epoch, GMM 100M x 128 sharded, fixed 64x64 map"),
    "c2": dict(per_gpu=70_000, total=None, d=784, side=20, k=10, name="SomClassifier-shaped epoch, GMM 70000 x 784, 20x20 map"),
    "c5": dict(per_gpu=None, total=5_000_000, d=4096, side=128, k=64, name="SomVQ epoch, GMM 5M x 4096 sharded, fixed 128x128 map (16384 neurons)"),
    "small": dict(per_gpu=200_000, total=None, d=64, side=16, k=16, name="GMM 200k x 64, 16x16 map"),
}
N_ITER_SCHEDULE = 200  # sigma follows the coarse phase of a 200-epoch fit


def sigma_at(epoch: int, m: int) -> float:
    """dbgsom/BaseSom.py:863-902 with default hyper-parameters (exponential decay, coarse phase)."""
    from math import exp, sqrt

    s0, s1 = 0.2 * sqrt(m), max(0.7, 0.05 * sqrt(m))
    return s1 + (s0 - s1) * exp(-0.02 * (epoch / 0.5))


def grid_hops(side: int) -> np.ndarray:
    ii, jj = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    p = np.stack([ii.ravel(), jj.ravel()], axis=1)
    return np.abs(p[:, None, :] - p[None, :, :]).sum(axis=2).astype(np.uint16)


def measured_peaks() -> dict:
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        p["_source"] = "measured"
        return p
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_source": "fallback"}


class ClockSampler:
    """nvidia-smi clock / throttle samples while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            if len(r) < 8:
                continue
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- data
def make_shard(torch, device, n, d, k, rank, chunk=1 << 20):
    """Synthetic Gaussian mixture shard generated on the device (centres N(0, 2^2), unit noise)."""
    gc = torch.Generator(device=device).manual_seed(20240)
    centers = torch.randn(k, d, device=device, generator=gc) * 2.0
    g = torch.Generator(device=device).manual_seed(1000 + rank)
    X = torch.empty((n, d), dtype=torch.float32, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        lab = torch.randint(0, k, (e - s,), device=device, generator=g)
        X[s:e] = centers[lab]
        X[s:e] += torch.randn(e - s, d, device=device, generator=g)
    return X


def cpu_reference_epoch_rate(wl, n_sample, steps, warmup):
    """Oracle (CPU restatement of the reference epoch) on a bounded row sample; returns a dict."""
    from threadpoolctl import threadpool_limits

    from oracle import som_oracle as O

    # torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core
    threadpool_limits(limits=os.cpu_count())
    d, side, k = wl["d"], wl["side"], wl["k"]
    m = side * side
    X = O.gmm(n_sample, d, k, seed=0)
    rng = np.random.default_rng(0)
    W = X[rng.choice(n_sample, m, replace=False)].astype(np.float64)
    hop = grid_hops(side).astype(np.float64)
    V = float(O.total_variance(X.astype(np.float64)))
    X64 = X.astype(np.float64)  # same dtype as W: sklearn's ArgKmin64 path, like the reference on float64 input
    times, t_lin, t_const = [], [], []
    for e in range(warmup + steps):
        t0 = time.perf_counter()
        dist, win = O.bmu(X64, W, 1)
        kk = O.sample_weights(dist, V)
        C, n = O.voronoi_centers(kk, X64, win, m, pack=True)
        E = O.quantization_errors(win, dist, m)
        t1 = time.perf_counter()
        H = O.neighborhood(hop, sigma_at(e, m))
        W = O.smooth(C, n, H)
        t2 = time.perf_counter()
        if e >= warmup:
            times.append(t2 - t0)
            t_lin.append(t1 - t0)
            t_const.append(t2 - t1)
    return dict(t_step=float(np.mean(times)), t_linear=float(np.mean(t_lin)), t_smooth=float(np.mean(t_const)),
                n_sample=n_sample, E_checksum=float(E.sum()))


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = min(wl["per_gpu"] or wl["total"], 100_000 if wl["side"] >= 32 else 70_000)
    r = cpu_reference_epoch_rate(wl, n_sample, args.steps, args.warmup)
    cores = os.cpu_count()
    value = n_sample / r["t_step"]
    sample = (f"{n_sample} x {wl['d']} rows of the workload, full {wl['side']}x{wl['side']} map; numpy float64 + "
              f"sklearn NearestNeighbors (ArgKmin64); BMU+update {r['t_linear']:.3f}s (linear in rows) + "
              f"smoothing GEMM {r['t_smooth']:.3f}s (independent of rows) per epoch")
    line = {
        "impl": "reference", "metric": "training samples/sec/epoch", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["t_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["name"], "rows_timed": n_sample, "d": wl["d"], "neurons": wl["side"] ** 2},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=args.out)


# ---------------------------------------------------------------------------------------------- ours
def run_ours(args, wl):
    import torch

    from dbgsom_b200 import _native as nat
    from dbgsom_b200.engine import DeviceEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    n_gpus = world

    d, side, k = wl["d"], wl["side"], wl["k"]
    m = side * side
    n_local = wl["per_gpu"] if wl["per_gpu"] else wl["total"] // n_gpus
    scaling = "weak" if wl["per_gpu"] else "strong"
    if args.rows:
        n_local = args.rows
    X = make_shard(torch, device, n_local, d, k, rank)

    eng = DeviceEngine(device=str(device), bmu_backend=args.backend, distributed=world > 1)
    eng.load_device_data(X)
    n_global = eng.n_samples_global
    rows = np.random.default_rng(0).choice(n_global, m, replace=False)
    eng.init_map_from_rows(rows, capacity=m)
    # all-pairs hop counts of the side x side grid: BFS on the device from the adjacency table (dbgsom_hops)
    from dbgsom_b200.topology import MapTopology

    eng.set_hops_from_topology(MapTopology.full_grid(side, side))

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    epoch = 0
    for _ in range(args.warmup):
        eng.epoch(sigma_at(epoch, m), True, False)
        epoch += 1

    # ------------------------------------------------------------------ timed region (device resident)
    sampler = ClockSampler(local)
    eng.enable_profiling(True)
    eng.bmu_stats_host(reset=True)
    launches0 = eng.launches
    barrier()
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    stats_acc = np.zeros(4)
    for _ in range(args.steps):
        out = eng.epoch(sigma_at(epoch, m), True, False)
        epoch += 1
    t1.record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = t0.elapsed_time(t1)
    launches = eng.launches - launches0
    bmu_stats = eng.bmu_stats_host()
    phases = eng.phase_times_ms()
    eng.enable_profiling(False)
    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = n_global * args.steps / (elapsed_ms * 1e-3)

    peaks = measured_peaks()
    be, n_pass = eng.last_backend
    mean = lambda xs: float(np.mean(xs)) if xs else None  # noqa: E731
    t_cand, t_acc = mean(phases.get("bmu_candidates")), mean(phases.get("accumulate"))
    flops = 2.0 * n_local * m * d
    roof = {
        "kernel": "bmu_cand_tensor_kernel" if be == nat.BMU_TENSOR else "bmu_cand_simt_kernel",
        "bound": "tensor", "unit": "TFLOP/s",
        "achieved": flops / (t_cand * 1e-3) / 1e12 if t_cand else None,
        "peak": peaks["bf16_tflops_sustained"], "peak_source": peaks["_source"] + " (sustained cuBLAS bf16)",
        "traffic": None, "ms_per_launch": t_cand, "algorithmic_flops_per_launch": flops,
        "mma_passes": n_pass if be == nat.BMU_TENSOR else 0,
    }
    roof["frac"] = roof["achieved"] / roof["peak"] if roof["achieved"] else None
    upd_bytes = n_local * (4.0 * d + 8.0) + 4.0 * (m * d + 3 * m)
    roof_upd = {
        "kernel": "dbgsom_accumulate (histogram + scan + scatter + segmented accumulate)", "bound": "hbm", "unit": "GB/s",
        "achieved": upd_bytes / (t_acc * 1e-3) / 1e9 if t_acc else None, "peak": peaks["hbm_gbs"],
        "peak_source": peaks["_source"] + " (copy)", "traffic": None, "ms_per_launch": t_acc,
        "algorithmic_bytes_per_launch": upd_bytes,
    }
    roof_upd["frac"] = roof_upd["achieved"] / roof_upd["peak"] if roof_upd["achieved"] else None
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            tr = json.load(f).get(args.workload, {})
        roof["traffic"] = tr.get(roof["kernel"])
        roof_upd["traffic"] = tr.get("accumulate_kernel")
    except Exception:
        pass

    # ------------------------------------------------------------------ end to end (host buffers)
    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty((n_local, d), dtype=torch.float32, pin_memory=True)
            host.copy_(X)
            torch.cuda.synchronize(device)
            k_e2e = max(1, min(args.steps, 3))
            eng.epoch_from_host(host, sigma_at(epoch, m), True, False)  # untimed: one-off stream / workspace setup
            epoch += 1
            barrier()
            w0 = time.perf_counter()
            step_ms = []
            for _ in range(k_e2e):
                ws = time.perf_counter()
                # H2D of this step's samples in row chunks on a copy stream; the fp16 shadow and the BMU
                # search of a chunk run while the next one is in flight; reads back error/count/change
                r = eng.epoch_from_host(host, sigma_at(epoch, m), True, False)
                w_host = eng.weights()                 # D2H of the prototypes (the step's result)
                epoch += 1
                step_ms.append(round(1e3 * (time.perf_counter() - ws), 2))
            barrier()
            dt = time.perf_counter() - w0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device=device)
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
                dt = float(t.item())
            e2e = {"value": n_global * k_e2e / dt, "unit": "samples/s", "steps": k_e2e, "step_ms": step_ms,
                   "h2d_bytes_per_step": int(n_local * d * 4), "d2h_bytes_per_step": int(w_host.nbytes + 8 * (3 * m + 5)),
                   "note": "per step: H2D of all samples from pinned host memory (chunked, overlapped with the fp16 shadow rebuild and the BMU search), update, smoothing, D2H of prototypes and per-neuron error"}
            del host
        except Exception as exc:  # e.g. pinned allocation refused
            e2e = {"value": None, "unit": "samples/s", "error": repr(exc)[:200]}

    cpu = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu:
        n_sample = min(n_local, 100_000 if side >= 32 else 70_000)
        r = cpu_reference_epoch_rate(wl, n_sample, 2, 1)
        full = n_local / (r["t_linear"] * n_local / n_sample + r["t_smooth"])
        cpu = {"value": full, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
               "sample": (f"{n_sample} rows of the workload, full map; BMU+update {r['t_linear']:.3f}s scaled linearly to "
                          f"{n_local} rows + smoothing GEMM {r['t_smooth']:.3f}s; measured on the sample: "
                          f"{n_sample / r['t_step']:.4g} samples/s")}

    if rank == 0:
        line = {
            "metric": "training samples/sec/epoch", "value": value, "unit": "samples/s", "n_gpus": n_gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f16x%d tensor cores + f64 re-score / f64 update / f64 smoothing" % n_pass
            if be == nat.BMU_TENSOR else "f32 search + f64 re-score / f64 update / f64 smoothing",
            "data": "synthetic",
            "config": {"workload": wl["name"], "rows_per_gpu": n_local, "rows_total": n_global, "d": d, "neurons": m,
                       "cache": "inputs (>= 10x L2) streamed from HBM every step, no flush needed"
                       if n_local * d * 4 > 1.2e9 else "inputs smaller than 10x L2",
                       "bmu_backend": {nat.BMU_TENSOR: f"tcgen05 fp16, {n_pass} MMA pass(es)", nat.BMU_SIMT: "fp32 CUDA cores"}[be],
                       "prototype_evolution": "real trajectory (reference semantics, packed rows), sigma schedule of a 200-epoch fit",
                       "parallelism": f"dp{n_gpus}"},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "roofline": roof, "roofline_update": roof_upd, "cpu_baseline": cpu,
            "phases_ms": {k_: mean(v) for k_, v in phases.items()},
            "bmu_rescore_per_epoch": {k_: v / args.steps for k_, v in bmu_stats.items()}, "last_change": out["change"],
        }
        print(json.dumps(line), file=args.out)
    eng.close()
    if world > 1:
        torch.distributed.destroy_process_group()


def _stdout_for_json_only():
    """Send everything libraries write to fd 1 (e.g. NCCL's version banner) to stderr; return a file for
    the one JSON line the driver parses."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--backend", default="auto", choices=["auto", "tensor", "tensor1", "simt"])
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (debugging)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    wl = WORKLOADS[args.workload]
    args.out = _stdout_for_json_only()
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)
    args.out.flush()


if __name__ == "__main__":
    main()
