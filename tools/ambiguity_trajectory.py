"""How many samples does a ONE-pass fp16 candidate search leave ambiguous over a whole fit?

    python tools/ambiguity_trajectory.py [rows] [d] [side] [n_iter] [backend]

Runs the config-3 workload (GMM rows x d, fixed side x side map) for n_iter epochs with the sigma
schedule of BaseSom._calculate_current_sigma (coarse phase = first half, then sigma_end) and prints,
per epoch, the re-score statistics and phase times of the chosen BMU back end.  Input to the choice
between one and three MMA passes per epoch.
"""
import os
import sys
from math import exp, sqrt

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from bench import make_shard  # noqa: E402
from dbgsom_b200.engine import DeviceEngine  # noqa: E402
from dbgsom_b200.topology import MapTopology  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    side = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    n_iter = int(sys.argv[4]) if len(sys.argv) > 4 else 200
    backend = sys.argv[5] if len(sys.argv) > 5 else "tensor1"
    m = side * side
    dev = torch.device("cuda", 0)
    X = make_shard(torch, dev, rows, d, 64, 0)
    eng = DeviceEngine(device="cuda:0", bmu_backend=backend)
    eng.load_device_data(X)
    eng.init_map_from_rows(np.random.default_rng(0).choice(rows, m, replace=False), capacity=m)
    eng.set_hops_from_topology(MapTopology.full_grid(side, side))
    eng.enable_profiling(True)
    s0, s1 = 0.2 * sqrt(m), max(0.7, 0.05 * sqrt(m))
    print("epoch sigma ambiguous flagged candidates rescans live cand_ms resolve_ms change")
    for e in range(n_iter):
        sigma = s1 + (s0 - s1) * exp(-0.02 * (e / 0.5)) if e <= 0.5 * n_iter else s1
        eng.bmu_stats_host(reset=True)
        out = eng.epoch(sigma, True, False)
        st = eng.bmu_stats_host()
        ph = eng.phase_times_ms()
        if e < 20 or e % 5 == 0:
            print(e, f"{sigma:.2f}", st["ambiguous"], st["flagged"], st["candidates"], st["full_rescans"],
                  int((out["counts"] > 0).sum()), "/".join(f"{v:.2f}" for v in ph['bmu_candidates']),
                  f"{np.mean(ph['bmu_resolve']):.2f}", f"{out['change']:.4g}", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
