"""What do the rows that overflow the candidate table look like on large collapsed maps (config-5 shape)?

    python tools/study_c5_groups.py <workload> <rows> <epochs>

Runs bench epochs, then for the state of the LAST search reports
  * groups of prototypes with bit-identical fp16 shadows (hi and lo): sizes, and the exact float64 distance of
    every member to its group's lowest-indexed member (the radius a near-duplicate exclusion would have to accept);
  * for a sample of rows: the number of prototypes J inside the acceptance window (2B around the exact minimum),
    the number of DISTINCT shadow groups among them, the diameter of J around the winner relative to the row's
    distance (triangle inequality: |d_a - d_b| <= ||w_a - w_b||, so diameter / d < 1e-6 proves the near-tie
    exemption of the parity gate), and 2B / d^2 (what the present near-tie proof needs below 1e-6).
"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, make_shard, sigma_at
from dbgsom_b200.engine import DeviceEngine
from dbgsom_b200.topology import MapTopology

wl = WORKLOADS[sys.argv[1]]
rows = int(sys.argv[2])
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 5
dev = torch.device("cuda", 0)
d, side = wl["d"], wl["side"]
m = side * side
X = make_shard(torch, dev, rows, d, wl["k"], 0)
eng = DeviceEngine(bmu_backend="tensor")
eng.load_device_data(X)
eng.init_map_from_rows(np.random.default_rng(0).choice(rows, m, replace=False), capacity=m)
eng.set_hops_from_topology(MapTopology.full_grid(side, side))
for e in range(epochs):
    out = eng.epoch(sigma_at(e, m), True, False)
    st = eng.bmu_stats_host(reset=True)
    print("epoch", e, {k: st[k] for k in ("ambiguous", "flagged", "full_rescans") if k in st},
          "live", int((out["counts"] > 0).sum()), flush=True)

W = eng.W[eng.cur ^ 1][:m, :d].clone()  # prototypes of the last search
mpad = eng.W16_hi.shape[0]
poc = eng.proto_of_col[:mpad].long()
hi = eng.W16_hi[:mpad].view(torch.int16)
lo = eng.W16_lo[:mpad].view(torch.int16)
valid = poc < m
key = torch.cat([hi, lo], dim=1)[valid]
pv = poc[valid]
uniq, inv, cnt = torch.unique(key, dim=0, return_inverse=True, return_counts=True)
res = {"prototypes": m, "shadow_groups": int(uniq.shape[0]), "in_groups_gt1": int(cnt[cnt > 1].sum()),
       "largest_groups": sorted(cnt.tolist(), reverse=True)[:8]}
# representative = lowest prototype index of the group
rep = torch.full((uniq.shape[0],), m, dtype=torch.long, device=dev)
rep.scatter_reduce_(0, inv, pv, reduce="amin")
rep_of_proto = torch.empty(m, dtype=torch.long, device=dev)
rep_of_proto[pv] = rep[inv]
grp_of_proto = torch.empty(m, dtype=torch.long, device=dev)
grp_of_proto[pv] = inv
dist_rep = (W - W[rep_of_proto]).pow(2).sum(1).sqrt()
res["member_to_rep_dist_q50_q99_max"] = torch.quantile(dist_rep[dist_rep > 0], torch.tensor([0.5, 0.99, 1.0], device=dev, dtype=torch.float64)).tolist() if (dist_rep > 0).any() else None
hi_only = torch.unique(hi[valid], dim=0).shape[0]
res["hi_shadow_groups"] = int(hi_only)
res["distinct_f64_rows"] = int(torch.unique(W, dim=0).shape[0])
um = (W - W.mean(0)).norm(dim=1)
res["u_norm_q50_max"] = [float(um.median()), float(um.max())]

ns = min(rows, 4096)
sel = torch.randperm(rows, device=dev)[:ns]
Xs = X[sel].double()
d2 = ((Xs * Xs).sum(1, keepdim=True) - 2 * Xs @ W.T + (W * W).sum(1)[None, :]).clamp_min(0)
dmin, win = d2.min(1)
xn = eng.xnorm16[sel].double()
wmax = eng.wmax.double()
coef = 0.0625 * 1.9073486e-6
xw = xn * wmax[0]
B = (xw * coef + 2.0 * 2.4e-7 * (xw + wmax[2])) / (eng.scale ** 2)
inb = d2 <= (dmin + 2 * B)[:, None]
nJ = inb.sum(1)
res["rows_sampled"] = ns
res["J_q50_q90_q99_max"] = torch.quantile(nJ.double(), torch.tensor([0.5, 0.9, 0.99, 1.0], device=dev, dtype=torch.float64)).tolist()
res["frac_J_gt8"] = float((nJ > 8).double().mean())
res["twoB_over_d2_q50_q90"] = torch.quantile(2 * B / dmin, torch.tensor([0.5, 0.9], device=dev, dtype=torch.float64)).tolist()
res["d_q50"] = float(dmin.sqrt().median())
# distinct shadow groups inside J, diameter of J around the winner
big = torch.nonzero(nJ > 8).flatten()[:512]
ng, diam_rel, gap_rel = [], [], []
for r in big.tolist():
    J = torch.nonzero(inb[r]).flatten()
    ng.append(int(torch.unique(grp_of_proto[J]).numel()))
    dd = (W[J] - W[win[r]]).pow(2).sum(1).sqrt().max()
    diam_rel.append(float(dd / dmin[r].sqrt()))
    s2 = torch.topk(d2[r], 2, largest=False).values.sqrt()
    gap_rel.append(float((s2[1] - s2[0]) / s2[0]))
if ng:
    a = np.array(ng, dtype=np.float64)
    res["flagged_rows_examined"] = len(ng)
    res["groups_in_J_q50_q90_max"] = [float(np.quantile(a, q)) for q in (0.5, 0.9, 1.0)]
    res["frac_flagged_rows_le8_groups"] = float((a <= 8).mean())
    dr = np.array(diam_rel)
    res["diamJ_over_d_q50_q90_max"] = [float(np.quantile(dr, q)) for q in (0.5, 0.9, 1.0)]
    res["frac_flagged_diam_below_1e-6"] = float((dr < 1e-6).mean())
    gr = np.array(gap_rel)
    res["true_gap_rel_q50_q90"] = [float(np.quantile(gr, q)) for q in (0.5, 0.9)]
    res["frac_flagged_true_gap_below_1e-6"] = float((gr < 1e-6).mean())

# ---- per-tile bounds (tile = 128 consecutive shadow columns) under three column orders
def tile_bounds(order, tile=128, floor=0.125):
    """B[rows, prototypes] with per-tile maxima of ||u|| and |wnorm| along `order` (prototype index per column)."""
    un = eng_un[order]
    wn = eng_wn[order]
    npad = (-len(order)) % tile
    if npad:
        un = torch.cat([un, un.new_zeros(npad)])
        wn = torch.cat([wn, wn.new_zeros(npad)])
    um_t = un.view(-1, tile).max(1).values.clamp_min(floor * un.max())
    wn_t = wn.view(-1, tile).max(1).values.clamp_min(floor * wn.max())
    um_c = um_t.repeat_interleave(tile)[: len(order)]
    wn_c = wn_t.repeat_interleave(tile)[: len(order)]
    um_p = torch.empty(m, dtype=torch.float64, device=dev)
    wn_p = torch.empty(m, dtype=torch.float64, device=dev)
    um_p[order] = um_c
    wn_p[order] = wn_c
    return um_p, wn_p

sc = float(eng.scale)
Wc = (W - W.mean(0)) * sc
eng_un = Wc.norm(dim=1)
vv = (W.mean(0) - eng.shift[:d].double()) * sc
eng_wn = (Wc.pow(2).sum(1) + 2 * (Wc @ vv)).abs()
acc = 2.0 * 2.4e-7
orders = {"scatter": poc[valid], "norm": torch.argsort(eng_un), "map": None}
try:
    from dbgsom_b200.engine import map_patch_order
    orders["map"] = torch.from_numpy(np.asarray(map_patch_order(MapTopology.full_grid(side, side).positions()), dtype=np.int64)).to(dev)
except Exception as ex:  # noqa: BLE001
    print("map order unavailable:", ex)
    orders.pop("map")
for name, order in orders.items():
    um_p, wn_p = tile_bounds(order)
    Bt = (xn[:, None] * um_p[None, :] * (coef + acc) + acc * wn_p[None, :]) / sc ** 2   # [rows, m]
    lo_ = d2 - Bt
    up_ = d2 + Bt
    U = up_.min(1).values
    U2 = torch.topk(up_, 2, largest=False).values[:, 1]
    L = lo_.min(1).values
    ncand = (lo_ <= U[:, None]).sum(1)
    proven = (U2 - L) <= 1e-6 * (dmin - 2 * B)
    over = ncand > 8
    res["tile_" + name] = {"frac_overflow": float(over.double().mean()), "frac_overflow_unproven": float((over & ~proven).double().mean()),
                           "cand_q50_q99_max": torch.quantile(ncand.double(), torch.tensor([0.5, 0.99, 1.0], device=dev, dtype=torch.float64)).tolist(),
                           "tile_umax_q10_q50_max": [float(torch.quantile(um_p, q)) for q in (0.1, 0.5, 1.0)]}
over0 = nJ > 8
proven0 = (torch.topk(d2, 2, largest=False).values[:, 1] - dmin + 2 * B) <= 1e-6 * (dmin - 2 * B)
res["global_bound"] = {"frac_overflow": float(over0.double().mean()), "frac_overflow_unproven": float((over0 & ~proven0).double().mean())}
print(json.dumps(res))
