"""CPU study (numpy, no GPU): what a ONE-pass fp16 candidate search leaves to a second stage on the
bench trajectory (config-3 shape at a reduced row count).  Emulates the shadows of dbgsom_prepare_x16 /
dbgsom_prepare_w, the one-pass score  wnorm_j - 2 xh.uh  and the per-row bound of bmu_tc.cu, and reports

  * candidates per row inside  min + 2B  (mean / quantiles / share above 8, 16, 32),
  * how local they are on the map: distinct 32-prototype patches (4 x 8 grid cells) per row,
  * for row tiles of 256 rows SORTED BY WINNER: the union of candidate patches -- the prototype columns a
    tile-selective second stage would have to visit.

    python tools/study_onepass.py [rows] [epochs]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import grid_hops, sigma_at  # noqa: E402
from oracle import som_oracle as O  # noqa: E402


def patch_of(side, ph, pw):
    ii, jj = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    return ((ii // ph) * (side // pw) + (jj // pw)).ravel()


def study(X, W, coef=0.25 * 2.0**-9, acc=2.4e-7, side=64, tag=""):
    n, d = X.shape
    m = W.shape[0]
    mean = X.mean(axis=0, dtype=np.float64)
    maxabs = np.abs(X).max()
    scale = 2.0 ** np.floor(np.log2(2.0**12 / (2.0 * maxabs)))
    xs = ((X - mean) * scale)
    wmean = W.mean(axis=0)
    u = (W - wmean) * scale
    v = (wmean - mean) * scale
    xh = xs.astype(np.float16).astype(np.float32)
    uh = u.astype(np.float16).astype(np.float32)
    wnorm = (np.einsum("ij,ij->i", u, u) + 2 * u @ v).astype(np.float32)
    xnorm = np.linalg.norm(xs, axis=1)
    umax = np.linalg.norm(u, axis=1).max()
    tau = 2 * (xnorm * umax * coef + acc * (xnorm * umax + np.abs(wnorm).max()))
    exact = (np.einsum("ij,ij->i", u, u) + 2 * u @ v)[None, :] - 2 * (xs.astype(np.float64) @ u.T)
    s = wnorm[None, :] - 2 * (xh @ uh.T)
    err = np.abs(s - exact).max(axis=1)
    smin = s.min(axis=1)
    inb = s <= (smin + tau)[:, None]
    c = inb.sum(axis=1)
    win = exact.argmin(axis=1)
    q = np.quantile(c, [0.5, 0.9, 0.99, 0.999])
    print(f"{tag} rows {n}: cand/row mean {c.mean():.1f} median {q[0]:.0f} p90 {q[1]:.0f} p99 {q[2]:.0f} p99.9 {q[3]:.0f} max {c.max()}"
          f" | >1: {np.mean(c > 1):.3f} >8: {np.mean(c > 8):.3f} >16: {np.mean(c > 16):.3f} >32: {np.mean(c > 32):.3f} >64: {np.mean(c > 64):.3f}")
    print(f"   bound tau/2 median {np.median(tau) / 2:.3g}, actual max |err| per row median {np.median(err):.3g} (ratio {np.median(tau / 2 / err):.1f}),"
          f" score spread (p50 of max-min) {np.median(s.max(axis=1) - smin):.3g}, live prototypes {np.unique(win).size}")
    for ph, pw in ((4, 8), (8, 16)):
        p = patch_of(side, ph, pw)
        npatch = p.max() + 1
        pm = np.zeros((n, npatch), dtype=bool)
        for k in range(npatch):
            pm[:, k] = inb[:, p == k].any(axis=1)
        per_row = pm.sum(axis=1)
        order = np.argsort(win, kind="stable")
        tiles = [order[i:i + 256] for i in range(0, n - 255, 256)]
        uni = np.array([pm[t].any(axis=0).sum() for t in tiles])
        amb = c > 1
        order_a = order[amb[order]]
        tiles_a = [order_a[i:i + 256] for i in range(0, order_a.size - 255, 256)]
        uni_a = np.array([pm[t].any(axis=0).sum() for t in tiles_a]) if tiles_a else np.array([0])
        print(f"   patches {ph}x{pw} ({ph * pw} protos, {npatch} total): per row mean {per_row.mean():.2f} p99 {np.quantile(per_row, 0.99):.0f};"
              f" union over winner-sorted 256-row tiles mean {uni.mean():.1f} p90 {np.quantile(uni, 0.9):.0f} max {uni.max()}"
              f" -> column share {uni.mean() / npatch:.3f}; ambiguous rows only ({amb.mean():.2f} of rows): union mean {uni_a.mean():.1f}"
              f" -> work share {amb.mean() * uni_a.mean() / npatch:.3f}")
    return c


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 13
    side, d = 64, 256
    m = side * side
    X = O.gmm(n, d, 64, seed=0)
    X64 = X.astype(np.float64)
    rng = np.random.default_rng(0)
    W = X[rng.choice(n, m, replace=False)].astype(np.float64)
    hop = grid_hops(side).astype(np.float64)
    V = float(O.total_variance(X64))
    for e in range(epochs):
        if e in (0, 1, 3, 6, 9, 12, 20, 40):
            for coef in (0.25 * 2.0**-9, 0.0625 * 2.0**-9, 2.0**-15):
                study(X, W, coef=coef, tag=f"epoch {e} coef {coef:.3g}")
        r = O.epoch_step(X64, W, hop, sigma_at(e, m), V, pack=True, bmu_fn=O.bmu_expansion)
        W = r["W_new"]
        print(f"epoch {e}: live {int((r['n'] > 0).sum())} change {r['change']:.4g}", flush=True)


if __name__ == "__main__":
    main()
