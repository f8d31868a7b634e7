"""GPU study (torch emulation, not a product path): how much of the prototype matrix a TILE-SELECTIVE second stage would
have to revisit if the first stage were ONE fp16 pass and the sample rows were kept sorted by their previous winner.

The engine runs the real bench trajectory (config-3 shape at a reduced row count); at chosen epochs the one-pass score
wnorm_j - 2 xh.uh and the per-row bound of bmu_tc.cu are emulated with torch and, with rows sorted by the winner of
`stale` epochs ago and prototypes cut into map patches (column tiles), the script reports

  * rows with >= 2 candidates inside the one-pass bound (they need the second stage),
  * per 256-row tile: how many column tiles hold an in-bound candidate of an ambiguous row (the (row tile, column tile)
    pairs the second stage computes with the split-fp16 passes), as a share of all pairs,
  * per 32-row warp: share of 32-column chunks with an in-bound score (how often the epilogue leaves its fast path).

    python tools/study_selective.py [rows] [epochs]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_shard, sigma_at  # noqa: E402
from dbgsom_b200.engine import DeviceEngine  # noqa: E402
from dbgsom_b200.topology import MapTopology  # noqa: E402


def patch_ids(side, ph, pw, dev):
    ii, jj = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    return torch.from_numpy(((ii // ph) * (side // pw) + (jj // pw)).ravel().astype(np.int64)).to(dev)


@torch.no_grad()
def analyse(X, W, prev_win, shift, scale, side, tag, coef=0.25 * 2.0**-9, acc=2.4e-7, prev_dirty=None):
    dev = X.device
    n, d = X.shape
    m = W.shape[0]
    wmean = W.mean(dim=0)
    u = (W - wmean) * scale
    v = (wmean - shift.double()) * scale
    uh = u.half().float()
    wnorm = ((u * u).sum(dim=1) + 2 * (u @ v)).float()
    umax = u.norm(dim=1).max().float()
    tiles = {"128 (8x16)": patch_ids(side, 8, 16, dev), "64 (8x8)": patch_ids(side, 8, 8, dev), "32 (4x8)": patch_ids(side, 4, 8, dev)}
    # sort rows by the column tile, then the index, of their previous winner
    key = tiles["128 (8x16)"][prev_win] * m + prev_win
    if prev_dirty is not None:  # minor key: how many candidates the row had in the previous search (0: 1, 1: 2-8, 2: 9-64, 3: more)
        bucket = (prev_dirty > 1).long() + (prev_dirty > 8).long() + (prev_dirty > 64).long()
        key = key * 4 + bucket
    order = torch.argsort(key)
    chunk = (1 << 17) if m <= 4096 else (1 << 14)
    ncand = torch.empty(n, dtype=torch.int32, device=dev)
    masks = {k: torch.zeros((n, int(t.max()) + 1), dtype=torch.bool, device=dev) for k, t in tiles.items()}
    for s in range(0, n, chunk):
        rows = order[s:s + chunk]
        xs = (X[rows] - shift) * scale
        xh = xs.half().float()
        xn = xs.norm(dim=1)
        sc = wnorm[None, :] - 2 * (xh @ uh.T)
        tau = 2 * (xn * umax * coef + acc * (xn * umax + wnorm.abs().max()))
        inb = sc <= (sc.min(dim=1).values + tau)[:, None]
        ncand[s:s + chunk] = inb.sum(dim=1)
        for k, t in tiles.items():
            mk = masks[k]
            for p in range(mk.shape[1]):
                mk[s:s + chunk, p] = inb[:, t == p].any(dim=1)
    amb = ncand >= 2
    q = torch.quantile(ncand.float()[: 1 << 20], torch.tensor([0.5, 0.9, 0.99], device=dev)).tolist()
    print(f"{tag}: ambiguous rows {amb.float().mean():.3f}, >8 cand {(ncand > 8).float().mean():.3f}, >64 {(ncand > 64).float().mean():.3f},"
          f" cand median {q[0]:.0f} p90 {q[1]:.0f} p99 {q[2]:.0f}")
    nt = n // 256 * 256
    for k, mk in masks.items():
        a = (mk & amb[:, None])[:nt].view(nt // 256, 256, -1).any(dim=1)  # [row tiles, column tiles]
        share = a.float().mean().item()
        per_tile = a.sum(dim=1).float()
        print(f"   column tiles of {k}: second-stage share of all (row tile, column tile) pairs {share:.3f}"
              f" (row tiles with any {(per_tile > 0).float().mean():.3f}, mean {per_tile.mean():.1f} p90 {torch.quantile(per_tile, 0.9):.0f} max {per_tile.max():.0f} of {a.shape[1]})")
    w32 = masks["32 (4x8)"][: n // 32 * 32].view(n // 32, 32, -1).any(dim=1)
    print(f"   warp level: {w32.float().mean():.3f} of (32-row warp, 32-column chunk) pairs hold an in-bound score")
    # the same with the rows in their ORIGINAL order (what the epilogue sees today)
    inv = torch.empty_like(order)
    inv[order] = torch.arange(n, device=dev)
    w32o = masks["32 (4x8)"][inv][: n // 32 * 32].view(n // 32, 32, -1).any(dim=1)
    print(f"   unsorted rows: {w32o.float().mean():.3f}")
    out = torch.empty_like(ncand)
    out[order] = ncand
    return out  # candidates per row, in sample order


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 13
    d = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    side = int(sys.argv[4]) if len(sys.argv) > 4 else 64
    m = side * side
    dev = torch.device("cuda", 0)
    X = make_shard(torch, dev, n, d, 64, 0)
    eng = DeviceEngine(device="cuda:0", bmu_backend="tensor")
    eng.load_device_data(X)
    eng.init_map_from_rows(np.random.default_rng(0).choice(n, m, replace=False), capacity=m)
    eng.set_hops_from_topology(MapTopology.full_grid(side, side))
    hist = []
    for e in range(epochs):
        W = eng.W[eng.cur][:m].clone()
        if e in (7, 8, 11, 12) and hist and os.environ.get("STUDY_DIRTY"):
            if e in (7, 11):
                dirty = analyse(X, W, hist[-1], eng.shift, eng.scale, side, f"epoch {e} (dirtiness source)")
            else:
                analyse(X, W, hist[-1], eng.shift, eng.scale, side, f"epoch {e}, sorted by winner only")
                analyse(X, W, hist[-1], eng.shift, eng.scale, side, f"epoch {e}, sorted by (winner, candidate-count bucket of epoch {e - 1})",
                        prev_dirty=dirty)
        elif e in (2, 4, 8, 12) and hist:
            analyse(X, W, hist[-1], eng.shift, eng.scale, side, f"epoch {e}, rows sorted by the winners of epoch {e - 1}")
            if len(hist) >= 4:
                analyse(X, W, hist[-4], eng.shift, eng.scale, side, f"epoch {e}, rows sorted by the winners of epoch {e - 4}")
            if e == 8:
                analyse(X, W, hist[-1], eng.shift, eng.scale, side, f"epoch {e}, half the bound", coef=0.125 * 2.0**-9)
        r = eng.epoch(sigma_at(e, m), True, False)
        hist.append(eng.idx.view(-1)[:n].long().clone())
        print(f"epoch {e}: live {int((r['counts'] > 0).sum())}", flush=True)


if __name__ == "__main__":
    main()
