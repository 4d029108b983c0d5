"""Vertex-addition stream on the GPU (mirror of train/graph/dynamic_graph_vertex.py:11-166).

The reference re-extracts `graph.subgraph(all active vertices)` and rebuilds a scipy id map on every
snapshot.  Here the parent graph is relabelled ONCE to arrival rank (so the active set is the id
prefix [0, n_active) and subgraph id == rank), and a snapshot is one count + scan + ballot-compaction
pass over the parent CSR (csrc/graph.cu: k_prefix_*), keeping the parent's edge-id order in every row.
"""
import numpy as np
import torch

from .. import utils
from .._native import Graph
from ..sampling import NID
from .dynamic_graph import DynamicGraph, labelled_mask
from .device_graph import DeviceGraph, edges_from


class ParentGraph:
    """The full static graph handed to DynamicGraphVertex (stands in for dgl.from_networkx(G) with
    ndata['feat'|'target'], pubmed.py:85-108): a directed edge list in edge-id order."""

    def __init__(self, src, dst, n_vertices):
        self.src = np.ascontiguousarray(src, dtype=np.int64)
        self.dst = np.ascontiguousarray(dst, dtype=np.int64)
        self.n = int(n_vertices)
        self.ndata = {}

    @classmethod
    def from_undirected(cls, u, v, n_vertices):
        """both directions per undirected edge, forward block then reverse block"""
        u, v = np.asarray(u, dtype=np.int64), np.asarray(v, dtype=np.int64)
        return cls(np.concatenate([u, v]), np.concatenate([v, u]), n_vertices)

    def __len__(self):
        return self.n

    def number_of_edges(self):
        return len(self.src)


class DynamicGraphVertex(DynamicGraph):
    def __init__(self, graph, snapshots, labelled_vertices, search_depth=2):
        super().__init__(graph, snapshots, labelled_vertices, search_depth)
        self.evolving_vertices = None
        self.vertex_per_snapshot = int(len(self.graph) / self.snapshots)

    def build(self, vertex_timestamps=None, ensure_labelled=None):
        if vertex_timestamps is None:
            raise NotImplementedError("random snapshots are not implemented in the reference either (:96-97)")
        items = list(vertex_timestamps.items()) if isinstance(vertex_timestamps, dict) else list(zip(*vertex_timestamps))
        items.sort(key=lambda kv: kv[1])                     # stable, like the reference (:50-53)
        ordered = np.array([kv[0] for kv in items], dtype=np.int64)
        V, vps = len(self.graph), self.vertex_per_snapshot
        if ensure_labelled is None:
            bounds = list(range(0, V, vps)) + [V]
            self.snapshot_vertices = [ordered[a:b].tolist() for a, b in zip(bounds[:-1], bounds[1:])]
        else:
            if not (0 <= ensure_labelled <= 1):
                raise AssertionError("ensure_labelled must be in [0, 1]")
            per = int(vps * ensure_labelled)
            lab = np.asarray(labelled_mask(self.labelled_vertices, ordered), dtype=bool)
            chunks, cur, cnt = [], [], 0
            for v, is_l in zip(ordered.tolist(), lab.tolist()):
                cur.append(v)
                cnt += int(is_l)
                if cnt == per:
                    chunks.append(cur)
                    cur, cnt = [], 0
            if cur:
                chunks.append(cur)
            self.snapshot_vertices = chunks
        self._bounds = np.cumsum([0] + [len(c) for c in self.snapshot_vertices])
        self._order = ordered
        self._rank = np.empty(V, dtype=np.int64)
        self._rank[ordered] = np.arange(V)

        # device side: relabel parent edges to arrival rank, canonical in-CSR via the insert kernels
        pg = self.graph
        feats, targets = pg.ndata["feat"], pg.ndata["target"]
        rank_dev = torch.as_tensor(self._rank, device="cuda")
        tmp = Graph(V, max(pg.number_of_edges(), 16))
        tmp.insert_vertices(V)
        tmp.insert_edges(rank_dev[torch.as_tensor(pg.src, device="cuda")], rank_dev[torch.as_tensor(pg.dst, device="cuda")],
                         symmetric=False)
        indptr, indices, eids = tmp.export_csr()
        self.sub_g = DeviceGraph(V, max(pg.number_of_edges(), 16), int(feats.shape[1]))
        self.sub_g.native.load_parent(indptr, indices, eids)
        del tmp
        order_dev = torch.as_tensor(ordered, device="cuda")
        f_dev = torch.as_tensor(feats).to("cuda", torch.float32)
        t_dev = torch.as_tensor(targets).to("cuda", torch.int64).reshape(-1)
        self.sub_g.features.write_permuted(f_dev, t_dev, order_dev)
        if self.sub_g._feat is not None:
            self.sub_g._feat.copy_(f_dev[order_dev])
        self.sub_g._target.copy_(t_dev[order_dev].reshape(-1, 1))
        self.sub_g._nid = order_dev

        self.evolving_vertices = list(self.snapshot_vertices[0])
        self.evolution_index = 1
        self._activate()

    def _activate(self):
        n = int(self._bounds[self.evolution_index])
        self.sub_g.native.set_active_prefix(n)
        self.subgraph_to_original_map = self._order[:n]
        m = utils.sparse1d(len(self.graph))
        m.vec = np.where(self._rank < n, self._rank, 0)
        self.original_to_subgraph_map = m

    def evolve(self):
        self.evolving_vertices += self.snapshot_vertices[self.evolution_index]
        self.evolution_index += 1
        self._activate()

    def get_added_vertices(self, delta=None):
        delta = 1 if delta is None else delta
        acc = set()
        for i in range(delta):
            acc = acc.union(set(self.snapshot_vertices[self.evolution_index - i - 1]))
        vertices = list(acc)
        return vertices, labelled_mask(self.labelled_vertices, vertices)

    def get_vertices_changed(self):
        return set(self.snapshot_vertices[self.evolution_index - 1]), self.search_depth

    def get_graph(self):
        return self.sub_g

    def __len__(self):
        return len(self.snapshot_vertices)

    def get_original_to_subgraph_map(self):
        return self.original_to_subgraph_map

    def get_subgraph_to_original_map(self):
        return self.subgraph_to_original_map
