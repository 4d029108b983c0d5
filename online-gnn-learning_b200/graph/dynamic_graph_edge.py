"""Edge-addition stream on the GPU (mirror of train/graph/dynamic_graph_edge.py:10-261).

Snapshot k is the slice [k*eps, (k+1)*eps) of the time-ordered (src, dst) stream.  Its unseen
endpoints become vertices (their feature/label rows are appended to the device store), then the
slice is inserted as src->dst followed by dst->src edges by ONE warp-cooperative insert launch
sequence (csrc/graph.cu) instead of DGL's COO concat + lazy CSC rebuild.  Precondition shared with
the reference (reddit.py:101-113): vertex ids are dense in first-appearance order.
"""
import numpy as np
import torch

from .dynamic_graph import DynamicGraph, IdentityMap, labelled_mask
from .device_graph import DeviceGraph, edges_from


class DynamicGraphEdge(DynamicGraph):
    def __init__(self, snapshots, labelled_vertices, search_depth=1):
        super().__init__(None, snapshots, labelled_vertices, search_depth)
        self.current_subgraph = None
        self.new_vertices = set()
        self.n_seen = 0
        self.edge_feats = None
        self.check_dense_ids = True

    # ---- build / evolve ---------------------------------------------------------------------
    def build(self, vertex_feats, targets, cuda=True, edge_timestamps=None, ensure_labelled=None, restrict=None,
              edge_feats=None, v_cap=None, keep_master=True):
        if edge_timestamps is None:
            raise NotImplementedError("random snapshots are not implemented in the reference either (:84-85)")
        if edge_feats is not None:
            raise NotImplementedError("edge features are inactive for every dataset (edge_feats: 0, SURVEY a22)")
        src, dst = edges_from(edge_timestamps)
        n_total = len(src)
        if restrict is not None and restrict < n_total:
            src, dst = src[:restrict], dst[:restrict]
        self.src, self.dst = src, dst
        self.vertex_feats, self.targets = vertex_feats, targets
        self.edges_per_snapshot = int(n_total / self.snapshots)
        n_vertices = int(vertex_feats.shape[0])
        n_feats = int(vertex_feats.shape[1])
        self.current_subgraph = DeviceGraph(v_cap or n_vertices, 2 * len(src), n_feats, keep_master=keep_master)
        self._apply_slice(0)
        self.evolution_index = 1
        self.subgraph_to_original_map = IdentityMap()
        self.original_to_subgraph_map = self.subgraph_to_original_map

    def _apply_slice(self, k):
        eps = self.edges_per_snapshot
        s, d = self.src[k * eps:(k + 1) * eps], self.dst[k * eps:(k + 1) * eps]
        v_old = self.n_seen
        v_new = max(v_old, int(max(s.max(), d.max())) + 1) if len(s) else v_old
        if self.check_dense_ids and len(s):
            ends = np.unique(np.concatenate([s, d]))
            fresh = ends[ends >= v_old]
            if len(fresh) != v_new - v_old:
                raise ValueError("edge stream is not relabelled to dense first-appearance vertex ids "
                                 "(precondition of the reference, reddit.py:101-113)")
        new = np.arange(v_old, v_new, dtype=np.int64)
        self.new_vertices = set(new.tolist())
        self.n_seen = v_new
        g = self.current_subgraph
        g.add_nodes(len(new), {"feat": self.vertex_feats[v_old:v_new], "target": self.targets[v_old:v_new]})
        g.add_edges(s, d, symmetric=True)          # forward edges then reverse edges (:214-215)

    def evolve(self):
        self._apply_slice(self.evolution_index)
        self.evolution_index += 1

    # ---- queries ----------------------------------------------------------------------------
    def get_added_vertices(self, delta=None):
        if delta is None:
            vertices = self.new_vertices
        else:
            eps = self.edges_per_snapshot
            lo, hi = (self.evolution_index - delta) * eps, self.evolution_index * eps
            vertices = np.unique(np.concatenate([self.src[lo:hi], self.dst[lo:hi]]))
        return vertices, labelled_mask(self.labelled_vertices, vertices)

    def get_vertices_changed(self):
        eps = self.edges_per_snapshot
        lo, hi = (self.evolution_index - 1) * eps, self.evolution_index * eps
        return set(np.unique(np.concatenate([self.src[lo:hi], self.dst[lo:hi]])).tolist()), self.search_depth

    def get_graph(self):
        return self.current_subgraph

    def __len__(self):
        return self.snapshots

    def get_original_to_subgraph_map(self):
        return self.original_to_subgraph_map

    def get_subgraph_to_original_map(self):
        return self.subgraph_to_original_map
