"""Worker of tests/test_gpu_multirank.py: run under torchrun with one rank per GPU.

Every rank trains on its contiguous shard through the distributed engine (NCCL all-reduce of the partial
sums, optionally row-sharded smoothing + all-gather); rank 0 repeats the same epochs on the full data with
a single-GPU engine and compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _datasets  # noqa: E402
from dbgsom_b200.engine import DeviceEngine  # noqa: E402
from dbgsom_b200.topology import MapTopology  # noqa: E402


def run(eng, X, y, c, rows, topo, sigmas):
    eng.load_data(X, y, c)
    eng.init_map_from_rows(rows, capacity=len(topo))
    eng.set_hops_from_topology(topo)
    outs = []
    for s in sigmas:
        W_in = eng.weights()
        r = eng.epoch(s, True, False)
        r["W_in"], r["winners"], r["W_out"] = W_in, eng.last_winners_host(), eng.weights()
        outs.append(r)
    st = eng.final_statistics(topo.positions(), topo.degrees())
    eng.final_winners()
    hist = eng.label_histogram(c)
    return outs, eng.weights(), st, hist


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n, d, c = 80000, 96, 5
    X, y = _datasets.gmm(n, d, c, 7, return_labels=True)
    topo = MapTopology.full_grid(32, 32)  # 1024 prototypes: the tcgen05 search incl. its selective form applies
    rows = np.random.default_rng(0).choice(n, len(topo), replace=False)
    sigmas = [2.0, 1.6, 1.2]
    per = -(-n // world)
    sl = slice(rank * per, min(n, (rank + 1) * per))
    backend = os.environ.get("MR_BACKEND", "tensor")
    eng = DeviceEngine(device=f"cuda:{local}", distributed=True, bmu_backend=backend)
    eng.resort_every = 1
    outs, W, st, hist = run(eng, X[sl], y[sl], c, rows, topo, sigmas)
    kinds = eng.bmu_stats_host()
    eng.close()
    # winners of all ranks, in sample order, for the oracle comparison on rank 0
    wins = []
    for r in outs:
        mine = torch.from_numpy(r["winners"]).to(f"cuda:{local}")
        parts = [torch.empty(min(n, (q + 1) * per) - q * per, dtype=mine.dtype, device=mine.device) for q in range(world)]
        dist.all_gather(parts, mine)
        wins.append(torch.cat(parts).cpu().numpy())
    ok = True
    if rank == 0:
        # (1) against the float64 oracle: every epoch starts from the device's prototypes; winners must agree outside
        # the 1e-6 near-tie gate, the update (teacher-forced on the exempt samples) to 1e-5
        from oracle import som_oracle as O

        hop = topo.hop_matrix()
        X64 = X.astype(np.float64)
        V = float(np.var(X64, axis=0).sum())
        for e_, (r, win, s) in enumerate(zip(outs, wins, sigmas)):
            _, ref_win, gap = O.bmu_with_gap(X64, r["W_in"])
            strict = gap >= 1e-6
            bad = np.flatnonzero((win != ref_win) & strict)
            assert strict.mean() > 0.5, f"epoch {e_}: only {strict.mean():.3f} of the samples outside the near-tie gate"
            assert bad.size == 0, f"epoch {e_}: {bad.size} winners differ outside the gate, rows {bad[:5]}, gaps {gap[bad[:5]]}"
            ref = O.epoch_step(X64, r["W_in"], hop, s, V, pack=True, winners=win)
            np.testing.assert_array_equal(r["counts"], ref["n"])
            np.testing.assert_allclose(r["error"], ref["E"], rtol=1e-5, atol=1e-6)
            assert np.abs(r["W_out"] - ref["W_new"]).max() / np.abs(ref["W_new"]).max() < 1e-5
        if backend == "tensor":
            assert kinds["selective_searches"] == len(sigmas) - 1, kinds  # the sharded epochs ran the FLAG + REFINE passes
        # (2) against one GPU holding all samples
        ref = DeviceEngine(device="cuda:0", distributed=False, bmu_backend=backend)
        r_outs, r_W, r_st, r_hist = run(ref, X, y, c, rows, topo, sigmas)
        ref.close()
        for a, b in zip(outs, r_outs):
            np.testing.assert_array_equal(a["counts"], b["counts"])
            np.testing.assert_allclose(a["error"], b["error"], rtol=1e-10)
            assert abs(a["change"] - b["change"]) <= 1e-9 * abs(b["change"])
        np.testing.assert_allclose(W, r_W, rtol=1e-10, atol=1e-12)
        for k in ("te_count", "qe_sum"):
            assert abs(st[k] - r_st[k]) <= 1e-10 * max(1.0, abs(r_st[k])), k
        _, _, gap_prev = O.bmu_with_gap(X64, outs[-1]["W_in"])
        near_prev = int((gap_prev < 1e-9).sum())
        assert np.abs(st["hits"] - r_st["hits"]).sum() <= 2 * near_prev
        if near_prev == 0:
            np.testing.assert_allclose(st["dens_sum"], r_st["dens_sum"], rtol=1e-9)
        else:
            assert abs(st["dens_sum"].sum() - r_st["dens_sum"].sum()) <= 1e-6 * abs(r_st["dens_sum"].sum())
        # label histogram of the final top-1 search: the sharded and the single-GPU prototypes agree to ~1e-12, so only
        # samples sitting between prototypes that are (nearly) exact copies may land in another cell
        _, _, gap_final = O.bmu_with_gap(X64, W)
        near = int((gap_final < 1e-9).sum())
        assert np.abs(hist[0] - r_hist[0]).sum() <= 2 * near, (np.abs(hist[0] - r_hist[0]).sum(), near)
        assert hist[0].sum() == r_hist[0].sum() == n
        if near == 0:
            np.testing.assert_array_equal(hist[1], r_hist[1])
        print("MULTIRANK_OK world=%d shard_min_work=%s" % (world, os.environ.get("DBGSOM_K3_SHARD_MIN_WORK")), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
