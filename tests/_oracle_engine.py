"""TEST-ONLY engine: drives the host estimator logic with the CPU oracle.

Implements the interface of `dbgsom_b200.engine.DeviceEngine` in numpy float64 on top of
`oracle.som_oracle`, so the `-m "not gpu"` suite can check growth rules, schedules and
fitted attributes against the reference's trajectory fixtures without a GPU.  It lives
under tests/ on purpose: the product package never imports the oracle.
"""
import numpy as np

from dbgsom_b200.hostmath import class_entropy
from dbgsom_b200.topology import HOP_INF
from oracle import som_oracle as O


class OracleEngine:
    sample_offset = 0

    def __init__(self, **_):
        self.epochs = 0

    def load_data(self, X, y, n_classes):
        self.X, self.y, self.n_classes = X, y, n_classes
        self.n_samples_global = X.shape[0]
        self.total_var = O.total_variance(X)
        std = np.std(X, axis=0, ddof=1)
        return {
            "n_samples": X.shape[0],
            "total_variance": self.total_var,
            "std_norm": np.linalg.norm(std),
        }

    def init_map_from_rows(self, rows, capacity):
        self.W = np.array(self.X[np.asarray(rows)], dtype=self.X.dtype)

    def set_hops(self, hop_u16):
        hop = hop_u16.astype(np.float64)
        hop[hop_u16 == HOP_INF] = np.inf
        self.hop = hop

    def epoch(self, sigma, pack_rows, entropy_error):
        out = O.epoch_step(self.X, self.W, self.hop, sigma, self.total_var, pack=pack_rows)
        self.W_prev = self.W
        self.W = out["W_new"]
        self.epochs += 1
        if entropy_error:
            m = self.W.shape[0]
            hist = np.bincount(out["winners"] * self.n_classes + self.y, minlength=m * self.n_classes)
            err = class_entropy(hist.reshape(m, self.n_classes))
        else:
            err = out["E"]
        return {"error": err, "counts": out["n"], "change": out["change"]}

    def apply_row_ops(self, ops, n_rows):
        W = np.zeros((n_rows, self.W.shape[1]))
        W[: self.W.shape[0]] = self.W
        for op in ops:
            w = 2 * W[op.a] - W[op.b]
            if op.c >= 0:
                w = (w + W[op.c]) / 2
            W[op.dst] = w
        self.W = W

    def weights(self):
        return np.array(self.W, dtype=np.float64)

    @property
    def n_previous_rows(self):
        return self.W_prev.shape[0]

    def keep_rows(self, alive):
        self.W = self.W[alive]

    def set_hops_from_topology(self, topo):
        self.set_hops(topo.hop_matrix_u16())

    def final_statistics(self, positions, degrees):
        """numpy restatement of dbgsom/BaseSom.py:116-119 (TE :924-953, QE :904-922, node statistics
        :181-211, u-matrix :320-337) on the oracle's BMU search."""
        from scipy.spatial.distance import cdist

        dist2, idx2 = self.bmu_train(2, previous=True)
        m = self.W_prev.shape[0]
        pos = np.asarray(positions, dtype=np.float64)
        sep = pos[idx2[:, 0]] - pos[idx2[:, 1]]
        te = float(np.count_nonzero(np.sqrt((sep * sep).sum(axis=1)) > 1.5))
        winners, dist = idx2[:, 0], dist2[:, 0]
        W = self.weights()
        total = degrees.sum()
        avg = (cdist(W, W) * degrees[None, :]).sum(axis=1) / total if total > 0 else np.full(len(W), np.nan)
        bw = avg.mean()
        hits = np.bincount(winners, minlength=m).astype(np.float64)
        kern = np.exp(-(dist**2) / (2 * bw**2)) / (bw * np.sqrt(2 * np.pi))
        dens = np.bincount(winners, weights=kern, minlength=m)
        te, qe = self.allreduce_scalars([te, float(dist.sum())])
        hits, dens = self.allreduce_arrays([hits, dens])
        return {"te_count": te, "qe_sum": qe, "hits": hits, "dens_sum": dens, "avg_dist": avg, "weights": W, "n_rows": m}

    def final_winners(self):
        _, idx = self.bmu_train(1)
        self._final = idx[:, 0].astype(np.int64)

    def winners_host(self):
        return self._final

    def label_histogram(self, n_classes):
        m = self.W.shape[0]
        flat = self._final * n_classes + self.y.astype(np.int64)
        counts = np.bincount(flat, minlength=m * n_classes).astype(np.float64)
        first = np.full(m * n_classes, np.iinfo(np.int64).max, dtype=np.int64)
        np.minimum.at(first, flat, np.arange(flat.size, dtype=np.int64) + self.sample_offset)
        (counts,) = self.allreduce_arrays([counts])
        (first,) = self.allreduce_arrays([first], op="min")
        return counts.reshape(m, n_classes), first.reshape(m, n_classes)

    def sparse_code(self, Xn, Wn, max_iter=1000):
        """The reference's own call (dbgsom/BaseSom.py:258-266) on the already normalised operands."""
        from sklearn.decomposition import SparseCoder

        coder = SparseCoder(dictionary=Wn, positive_code=True, transform_alpha=0, transform_algorithm="lasso_lars")
        return coder.transform(Xn)

    def bmu_train(self, n_bmu, previous=False):
        return self.bmu(self.X, self.W_prev if previous else self.W, n_bmu)

    def bmu(self, X, W, n_bmu):
        dist, idx = O.bmu(X, W, n_bmu)
        return dist.reshape(X.shape[0], n_bmu), idx.reshape(X.shape[0], n_bmu)

    def allreduce_scalars(self, vals):
        return list(vals)

    def allreduce_arrays(self, arrs, op="sum"):
        return list(arrs)

    def close(self):
        pass


class ShardedOracleEngine(OracleEngine):
    """TEST-ONLY: the N > 1 host path on CPU.  Every rank holds a contiguous shard of the samples;
    per epoch only the per-neuron partial sums [Sk | sk | n | E] are all-reduced (gloo), then every
    rank smooths redundantly -- the same protocol the CUDA engine runs over NCCL."""

    def __init__(self, comm, **kw):
        super().__init__(**kw)
        self.comm = comm

    def _allreduce(self, arr, op="sum"):
        import torch

        t = torch.from_numpy(np.ascontiguousarray(arr))
        self.comm.allreduce_(t, op)
        return t.numpy()

    def load_data(self, X, y, n_classes):
        self.X, self.y, self.n_classes = X, y, n_classes
        counts = np.zeros(self.comm.world, dtype=np.int64)
        counts[self.comm.rank] = X.shape[0]
        counts = self._allreduce(counts)
        self.sample_offset = int(counts[: self.comm.rank].sum())
        self.n_samples_global = int(counts.sum())
        n = self.n_samples_global
        s1 = self._allreduce(X.astype(np.float64).sum(axis=0))
        mean = s1 / n
        ss = self._allreduce(((X.astype(np.float64) - mean) ** 2).sum(axis=0))
        self.total_var = float((ss / n).sum())
        return {"n_samples": n, "total_variance": self.total_var, "std_norm": float(np.sqrt((ss / (n - 1)).sum()))}

    def init_map_from_rows(self, rows, capacity):
        rows = np.asarray(rows)
        local = rows - self.sample_offset
        own = (local >= 0) & (local < self.X.shape[0])
        W = np.zeros((len(rows), self.X.shape[1]))
        W[own] = self.X[local[own]]
        self.W = self._allreduce(W)

    def epoch(self, sigma, pack_rows, entropy_error):
        M, D = self.W.shape
        dist, win = O.bmu(self.X, self.W, 1)
        k = O.sample_weights(dist, self.total_var)
        Sk, sk, n = O.voronoi_sums(k, self.X, win, M)
        E = O.quantization_errors(win, dist, M)
        buf = self._allreduce(np.concatenate([Sk.ravel(), sk, n, E]))
        Sk, sk, n, E = buf[: M * D].reshape(M, D), buf[M * D : M * D + M], buf[M * D + M : M * D + 2 * M], buf[M * D + 2 * M :]
        live = np.flatnonzero(n > 0)
        C = np.zeros((M, D))
        if pack_rows:
            C[: live.size] = Sk[live] / sk[live][:, None]
        else:
            C[live] = Sk[live] / sk[live][:, None]
        W_new = O.smooth(C, n, O.neighborhood(self.hop, sigma))
        change = float(np.sum(np.linalg.norm(self.W - W_new, axis=1)))
        self.W_prev, self.W = self.W, W_new
        if entropy_error:
            hist = np.bincount(win * self.n_classes + self.y, minlength=M * self.n_classes).astype(np.float64)
            E = class_entropy(self._allreduce(hist).reshape(M, self.n_classes))
        return {"error": E, "counts": n, "change": change}

    def allreduce_scalars(self, vals):
        return list(self._allreduce(np.asarray(vals, dtype=np.float64)))

    def allreduce_arrays(self, arrs, op="sum"):
        return [self._allreduce(a, op) for a in arrs]
