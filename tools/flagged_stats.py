"""How many prototypes lie inside the acceptance window of the three-pass tensor search?

    python tools/flagged_stats.py <workload> <rows> <epochs>

Runs a few bench epochs, then scores a sample of rows against all prototypes in float64 (torch) and
prints quantiles of the number of prototypes within 2B and 4B of the exact minimum, B = the per-sample
bound of the tensor front end (csrc/common.cuh tensor_score_bound) in squared-distance units.
"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, make_shard, sigma_at
from dbgsom_b200.engine import DeviceEngine
from dbgsom_b200.topology import MapTopology

wl = WORKLOADS[sys.argv[1]]
rows = int(sys.argv[2])
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
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
    st = eng.bmu_stats_host()
    print("epoch", e, st, "live", int((out["counts"] > 0).sum()), flush=True)
# state of the LAST search: prototypes before the last update = W[cur ^ 1]
W = eng.W[eng.cur ^ 1][:m, :d]
ns = min(rows, 2048)
sel = torch.randperm(rows, device=dev)[:ns]
Xs = X[sel].double()
d2 = (Xs * Xs).sum(1, keepdim=True) - 2 * Xs @ W.T + (W * W).sum(1)[None, :]
dmin = d2.min(1).values
xn = eng.xnorm16[sel].double()
wmax = eng.wmax.double()
coef = 0.0625 * 1.9073486e-6
xw = xn * wmax[0]
B = (xw * coef + 2.4e-7 * (xw + wmax[2])) / (eng.scale ** 2)
res = {}
for mult in (2, 4):
    cnt = (d2 <= (dmin + mult * B)[:, None]).sum(1).float()
    q = torch.quantile(cnt, torch.tensor([0.5, 0.9, 0.99, 1.0], device=dev)).tolist()
    res[f"within_{mult}B"] = {"q50,q90,q99,max": q, "frac_gt8": float((cnt > 8).float().mean()),
                              "frac_gt32": float((cnt > 32).float().mean()), "frac_gt128": float((cnt > 128).float().mean())}
res["rel_bound_4B_over_d2_median"] = float((4 * B / dmin).median())
uniq = torch.unique(W, dim=0).shape[0]
res["distinct_prototype_rows"] = int(uniq)
print(json.dumps(res))
