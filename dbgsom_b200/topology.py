"""Host-side map topology and growth rules of the directed batch growing SOM.

The reference keeps the map in a NetworkX graph and edits it in place
(`dbgsom/BaseSom.py:563-861`).  Growth decisions stay on the host in this build too
(BASELINE.json north star), but the prototypes live on the GPU, so the rules here never
touch weight vectors: a growth pass consumes the per-neuron error vector that the device
epoch produced and emits a list of *row operations* (`RowOp`) which the device applies to
its float64 prototype matrix in order.  The topology itself is an insertion-ordered
array structure (positions, adjacency lists in edge-insertion order) from which the
fitted `som_` NetworkX graph is exported at the end of `fit`.

Order matters for parity with the reference and is preserved deliberately:
  * prototype rows are in node-insertion order (new rows are always appended);
  * adjacency lists are in edge-insertion order (`nbr1, nbr2 = som_.adj[bo]`,
    `dbgsom/BaseSom.py:677`, depends on it);
  * error redistribution walks nodes in insertion order and mutates in place
    (`dbgsom/BaseSom.py:569-586`);
  * insertion candidates are visited by `np.argsort(-errors)` and the pass stops at the
    first node that is below the threshold or interior (`dbgsom/BaseSom.py:593-614`).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

Pos = tuple  # (x, y) integer grid position

# neighbour probe orders used by the reference
_CONNECT_ORDER = ((0, 1), (0, -1), (-1, 0), (1, 0))  # dbgsom/BaseSom.py:854-859
_FREE_SLOT_ORDER = ((0, 1), (0, -1), (1, 0), (-1, 0))  # dbgsom/BaseSom.py:628-633

HOP_INF = 0xFFFF  # "unreachable" in the uint16 hop matrix handed to the device


@dataclass(frozen=True)
class RowOp:
    """W[dst] = 2*W[a] - W[b]            when c < 0   (dbgsom/BaseSom.py:641-644, :705-726, :835-837)
    W[dst] = ((2*W[a] - W[b]) + W[c])/2  when c >= 0  (dbgsom/BaseSom.py:824-827)"""

    dst: int
    a: int
    b: int
    c: int = -1


class MapTopology:
    """Insertion-ordered 4-connected grid graph with per-node error and creation epoch."""

    def __init__(self) -> None:
        self.pos: list[Pos] = []
        self.index: dict[Pos, int] = {}
        self.adj: list[list[int]] = []
        self.edges: list[tuple[int, int]] = []
        self.epoch_created: list[int] = []
        self.error = np.zeros(0, dtype=np.float64)

    # ------------------------------------------------------------------ construction
    @classmethod
    def initial_square(cls) -> "MapTopology":
        """Four neurons in a square -- `_create_som`, dbgsom/BaseSom.py:425-442."""
        t = cls()
        for p in ((0, 0), (0, 1), (1, 0), (1, 1)):
            t._append(p, 0)
        for a, b in (((0, 0), (0, 1)), ((0, 0), (1, 0)), ((1, 0), (1, 1)), ((0, 1), (1, 1))):
            t._connect(t.index[a], t.index[b])
        return t

    @classmethod
    def full_grid(cls, gx: int, gy: int) -> "MapTopology":
        """Fixed gx x gy map in the node and edge order of `networkx.grid_2d_graph(gx, gy)`.

        The reference has no fixed-size mode; benchmark configs 3/4 express a fixed 64x64
        map through the step-level recipe, which this mirrors.
        """
        t = cls()
        for i in range(gx):
            for j in range(gy):
                t._append((i, j), 0)
        # grid_2d_graph adds all "down" edges (i -> i+1) first, then all "right" edges
        for i in range(gx - 1):
            for j in range(gy):
                t._connect(t.index[(i, j)], t.index[(i + 1, j)])
        for i in range(gx):
            for j in range(gy - 1):
                t._connect(t.index[(i, j)], t.index[(i, j + 1)])
        return t

    def _append(self, p: Pos, epoch: int) -> int:
        i = len(self.pos)
        self.pos.append(p)
        self.index[p] = i
        self.adj.append([])
        self.epoch_created.append(epoch)
        self.error = np.append(self.error, 0.0)
        return i

    def _connect(self, a: int, b: int) -> None:
        if b in self.adj[a]:
            return  # re-adding an existing edge does not reorder adjacency
        self.adj[a].append(b)
        self.adj[b].append(a)
        self.edges.append((a, b))

    def __len__(self) -> int:
        return len(self.pos)

    def degree(self, i: int) -> int:
        return len(self.adj[i])

    # ------------------------------------------------------------------ growth
    def place(self, p: Pos, epoch: int) -> int:
        """`_add_node_to_graph` + `_add_new_connections`, dbgsom/BaseSom.py:840-861.

        A position that already holds a neuron is overwritten in place (same row): its
        error is reset and its creation epoch updated, its edges are kept.
        """
        i = self.index.get(p)
        if i is None:
            i = self._append(p, epoch)
        else:
            self.epoch_created[i] = epoch
        self.error[i] = 0.0
        x, y = p
        for dx, dy in _CONNECT_ORDER:
            j = self.index.get((x + dx, y + dy))
            if j is not None:
                self._connect(i, j)
        return i

    def distribute_errors(self, threshold: float) -> None:
        """`_distribute_errors`, dbgsom/BaseSom.py:563-586 (in place, insertion order)."""
        err = self.error
        for i, nbrs in enumerate(self.adj):
            if len(nbrs) < 4:
                continue
            e = err[i]
            if e > threshold:
                boundary = [j for j in nbrs if len(self.adj[j]) < 4]
                if boundary:
                    share = 0.5 * e / len(boundary)
                    for j in boundary:
                        err[j] += share
                err[i] = err[i] / 2

    def grow(self, threshold: float, epoch: int) -> list[RowOp]:
        """`_add_new_neurons`, dbgsom/BaseSom.py:588-614.  Returns the prototype row ops."""
        snapshot = self.error.copy()
        order = np.argsort(-snapshot)
        ops: list[RowOp] = []
        for i in order:
            i = int(i)
            deg = len(self.adj[i])
            if snapshot[i] > threshold and deg < 4:
                if deg == 3:
                    p, a, b, c = self._one_free(i)
                elif deg == 2:
                    p, a, b, c = self._two_free(i)
                elif deg == 1:
                    p, a, b, c = self._three_free(i)
                else:
                    continue
                dst = self.place(p, epoch)
                ops.append(RowOp(dst, a, b, c))
            else:
                break
        return ops

    # Each rule returns (new position, a, b, c) meaning W_new = 2 W[a] - W[b] (and, if
    # c >= 0, the mean of that with W[c]).
    def _one_free(self, i: int):
        """One free slot -- `_insert_neuron_1p`, dbgsom/BaseSom.py:616-646."""
        x, y = self.pos[i]
        taken = {self.pos[j] for j in self.adj[i]}
        free = None
        for dx, dy in _FREE_SLOT_ORDER:
            q = (x + dx, y + dy)
            if q not in taken:
                free = q  # no early exit in the reference: the last free slot wins
        opposite = self.index[(2 * x - free[0], 2 * y - free[1])]
        return free, i, opposite, -1

    def _two_free(self, i: int):
        """Two free slots -- `_insert_neuron_2p`, dbgsom/BaseSom.py:648-728."""
        n1, n2 = self.adj[i]
        (x, y), (x1, y1), (x2, y2) = self.pos[i], self.pos[n1], self.pos[n2]
        if self.error[n1] > self.error[n2]:
            p, b = (2 * x - x2, 2 * y - y2), n2
        else:
            p, b = (2 * x - x1, 2 * y - y1), n1
        if x1 == x2 or y1 == y2:  # the two neighbours face each other
            if x1 == x2:
                p, b = (x + 1, y), n2
            else:
                p, b = (x, y + 1), n1
        return p, i, b, -1

    def _three_free(self, i: int):
        """Three free slots -- `_insert_neuron_3p` and cases a/b/c, dbgsom/BaseSom.py:730-838."""
        x, y = self.pos[i]
        n1 = self.adj[i][0]
        diagonals = {(x + 1, y + 1), (x + 1, y - 1), (x - 1, y + 1), (x - 1, y - 1)}
        # same set expression as the reference so that tie cases enumerate identically
        around = list(diagonals.intersection(set(self.pos[j] for j in self.adj[n1])))
        err = self.error
        if len(around) == 0:
            return self._extend(i, n1)
        if len(around) == 1:
            return self._side(i, n1, self.index[around[0]])
        n2, n3 = self.index[around[0]], self.index[around[1]]
        if err[n1] > err[n2] and err[n1] > err[n3]:
            return self._extend(i, n1)
        if err[n2] > err[n3]:
            return self._side(i, n1, n2)
        return self._side(i, n1, n3)

    def _extend(self, i: int, n1: int):
        """Straight continuation away from the only neighbour -- `_3p_case_c`, :831-838."""
        (x, y), (x1, y1) = self.pos[i], self.pos[n1]
        return (2 * x - x1, 2 * y - y1), i, n1, -1

    def _side(self, i: int, n1: int, n2: int):
        """`_3p_case_b`, dbgsom/BaseSom.py:813-829."""
        if self.error[n1] > self.error[n2]:
            return self._extend(i, n1)
        (x, y), (x1, y1), (x2, y2) = self.pos[i], self.pos[n1], self.pos[n2]
        return (x2 + x - x1, y2 + y - y1), i, n1, n2

    # ------------------------------------------------------------------ derived data
    def hop_matrix(self) -> np.ndarray:
        """All-pairs hop counts (float64, inf when unreachable), node order.

        Equals `nx.floyd_warshall_numpy(som_)` (dbgsom/BaseSom.py:401) on the unweighted
        graph but costs O(M * (M + E)) breadth-first searches instead of O(M^3).
        """
        from scipy.sparse import csr_matrix
        from scipy.sparse.csgraph import shortest_path

        m = len(self.pos)
        if m == 0:
            return np.zeros((0, 0))
        if not self.edges:
            d = np.full((m, m), np.inf)
            np.fill_diagonal(d, 0.0)
            return d
        e = np.asarray(self.edges, dtype=np.int64)
        g = csr_matrix((np.ones(len(e)), (e[:, 0], e[:, 1])), shape=(m, m))
        return shortest_path(g, method="D", directed=False, unweighted=True)

    def hop_matrix_u16(self) -> np.ndarray:
        d = self.hop_matrix()
        out = np.full(d.shape, HOP_INF, dtype=np.uint16)
        finite = np.isfinite(d)
        if finite.any() and d[finite].max() >= HOP_INF:
            raise ValueError("map diameter exceeds the uint16 hop encoding")
        out[finite] = d[finite].astype(np.uint16)
        return out

    def positions(self) -> np.ndarray:
        return np.asarray(self.pos, dtype=np.int64).reshape(-1, 2)

    def adjacency_table(self) -> np.ndarray:
        """int32 [M, 4]: neighbour indices per node (edge-insertion order), -1 = free slot.
        Input of the device hop-matrix kernel (`dbgsom_hops`)."""
        out = np.full((len(self.pos), 4), -1, dtype=np.int32)
        for i, nbrs in enumerate(self.adj):
            out[i, : len(nbrs)] = nbrs
        return out

    def degrees(self) -> np.ndarray:
        return np.array([len(a) for a in self.adj], dtype=np.float64)

    def to_networkx(self, node_attrs: dict[str, list] | None = None):
        """Export as the `som_` graph: same node order and the same adjacency order."""
        import networkx as nx

        g = nx.Graph()
        for i, p in enumerate(self.pos):
            g.add_node(p, epoch_created=self.epoch_created[i], error=self.error[i])
        g.add_edges_from((self.pos[a], self.pos[b]) for a, b in self.edges)
        if node_attrs:
            for name, values in node_attrs.items():
                for p, v in zip(self.pos, values):
                    g.nodes[p][name] = v
        return g

    def without(self, dead: np.ndarray) -> "MapTopology":
        """Copy with the listed nodes removed (order of the rest kept) -- the topology part
        of `_delete_dead_neurons_from_graph`, dbgsom/BaseSom.py:223-235."""
        dead_set = set(int(i) for i in dead)
        keep = [i for i in range(len(self.pos)) if i not in dead_set]
        remap = {old: new for new, old in enumerate(keep)}
        t = MapTopology()
        for old in keep:
            t._append(self.pos[old], self.epoch_created[old])
        t.error = self.error[keep].copy()
        for a, b in self.edges:
            if a in remap and b in remap:
                t._connect(remap[a], remap[b])
        return t
