"""Canonical streaming graph -- oracle only (numpy).

Restates what the reference asks DGL to do:
  * edge streams   train/graph/dynamic_graph_edge.py:52-82 (build) and :190-218 (evolve):
    snapshot k is the slice [k*eps, (k+1)*eps) of the time-ordered (src, dst) list;
    previously unseen endpoints become vertices (sorted-unique order), then the slice
    is appended as src->dst edges followed by dst->src edges (:71-72, :214-215).
    Edge id == insertion order (DGL add_edges appends).
  * vertex streams train/graph/dynamic_graph_vertex.py:82-94,132-141: the active set is a
    prefix of the timestamp-sorted vertex list, the snapshot graph is the subgraph
    induced by it, subgraph node i <-> evolving_vertices[i], induced edges keep the
    parent's edge-id order [recalled DGL 0.5 behaviour].
The in-neighbour list of v is [src(e) for e ascending if dst(e) == v] (what a stable
COO->CSC conversion yields); degrees are the list lengths.
"""
import numpy as np


def in_csr(src, dst, n_vertices):
    """Stable COO -> in-edge CSR.  Returns (indptr[int64 V+1], indices[int64 E], eids[int64 E])."""
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    order = np.argsort(dst, kind="stable")
    deg = np.bincount(dst, minlength=n_vertices).astype(np.int64)
    indptr = np.zeros(n_vertices + 1, dtype=np.int64)
    np.cumsum(deg, out=indptr[1:])
    return indptr, src[order], order.astype(np.int64)


class EdgeStreamOracle:
    """Edge-addition stream (dynamic_graph_edge.py).  ``src``/``dst`` are the full
    time-ordered stream; vertex ids must already be dense in first-appearance order
    (precondition of the reference: reddit.py:101-113 relabels)."""

    def __init__(self, src, dst, snapshots):
        self.src = np.asarray(src, dtype=np.int64)
        self.dst = np.asarray(dst, dtype=np.int64)
        self.snapshots = int(snapshots)
        self.edges_per_snapshot = int(len(self.src) / self.snapshots)      # :59
        self.e_src = np.zeros(0, dtype=np.int64)   # directed edge log, edge-id order
        self.e_dst = np.zeros(0, dtype=np.int64)
        self.seen = set()
        self.new_vertices = []
        self.n_vertices = 0
        self.evolution_index = 0
        self._apply(0)
        self.evolution_index = 1

    def _apply(self, k):
        eps = self.edges_per_snapshot
        s, d = self.src[k * eps:(k + 1) * eps], self.dst[k * eps:(k + 1) * eps]
        uniq = np.unique(np.concatenate([s, d]))            # sorted unique (:61, :195)
        new = [int(x) for x in uniq if int(x) not in self.seen]   # :200-204
        self.seen.update(new)
        self.new_vertices = new
        self.n_vertices += len(new)
        self.e_src = np.concatenate([self.e_src, s, d])     # forward then reverse (:214-215)
        self.e_dst = np.concatenate([self.e_dst, d, s])

    def evolve(self):
        self._apply(self.evolution_index)
        self.evolution_index += 1

    def csr(self):
        return in_csr(self.e_src, self.e_dst, self.n_vertices)

    def touched(self, delta):
        """get_added_vertices_pandas(delta) (:109-124): unique endpoints of the last delta slices."""
        eps = self.edges_per_snapshot
        lo, hi = (self.evolution_index - delta) * eps, self.evolution_index * eps
        return np.unique(np.concatenate([self.src[lo:hi], self.dst[lo:hi]]))


class VertexStreamOracle:
    """Vertex-addition stream (dynamic_graph_vertex.py).  Parent graph = directed edge
    list (p_src, p_dst) in edge-id order over V vertices; ``order`` = vertex ids sorted by
    timestamp (stable, :50-53)."""

    def __init__(self, p_src, p_dst, n_vertices, order, snapshots):
        self.p_src = np.asarray(p_src, dtype=np.int64)
        self.p_dst = np.asarray(p_dst, dtype=np.int64)
        self.V = int(n_vertices)
        self.order = np.asarray(order, dtype=np.int64)
        self.vertex_per_snapshot = int(self.V / int(snapshots))          # :29
        vps = self.vertex_per_snapshot
        self.chunks = [self.order[i:i + vps] for i in range(0, self.V, vps)]   # :56-57
        self.evolution_index = 1
        self.rank = np.full(self.V, -1, dtype=np.int64)    # original id -> subgraph id
        self.rank[self.order] = np.arange(self.V)

    def __len__(self):
        return len(self.chunks)

    def n_active(self):
        return int(sum(len(c) for c in self.chunks[:self.evolution_index]))

    def evolve(self):
        self.evolution_index += 1

    def subgraph_to_original(self):
        return self.order[:self.n_active()]

    def csr(self):
        """In-CSR of the induced subgraph in subgraph ids; eids are PARENT edge ids."""
        n = self.n_active()
        rs, rd = self.rank[self.p_src], self.rank[self.p_dst]
        keep = np.nonzero((rs < n) & (rd < n))[0]            # parent edge-id order
        indptr, indices, pos = in_csr(rs[keep], rd[keep], n)
        return indptr, indices, keep[pos]
