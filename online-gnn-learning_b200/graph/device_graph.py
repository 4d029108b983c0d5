"""The graph object the trainers see (what `DynamicGraph.get_graph()` returns): a DGL-graph
look-alike whose storage is the on-GPU streaming CSR + feature store of the C library."""
import numpy as np
import torch

from .. import config
from .._native import Graph, Features, Plan
from ..sampling import NID


class DeviceGraph:
    def __init__(self, v_cap, e_cap_directed, n_feats, mode=None, keep_master=True):
        self.mode = config.mode() if mode is None else mode
        self.native = Graph(v_cap, max(int(e_cap_directed), 16))
        self.features = Features(v_cap, n_feats, self.mode)
        self.v_cap, self.n_feats = int(v_cap), int(n_feats)
        self._feat = torch.zeros(v_cap, n_feats, dtype=torch.float32, device="cuda") if keep_master else None
        self._target = torch.full((v_cap, 1), -1, dtype=torch.int64, device="cuda")
        self._nid = None
        self.edata = {}
        self._splan = {}

    # ---- DGL-like surface ---------------------------------------------------
    @property
    def ndata(self):
        n = self.number_of_nodes()
        d = {"target": self._target[:n]}
        if self._feat is not None:
            d["feat"] = self._feat[:n]
        if self._nid is not None:
            d[NID] = self._nid[:n]
        return d

    def number_of_nodes(self):
        return self.native.num_vertices

    def number_of_edges(self):
        return self.native.num_edges

    def __len__(self):
        return self.number_of_nodes()

    def nodes(self):
        return torch.arange(self.number_of_nodes(), device="cuda")

    def in_degrees(self):
        return self.native.degrees()

    def to(self, device):
        return self

    def predecessors(self, v):
        """source vertices of the in-edges of `v`, in edge-id order (DGL `g.predecessors`; used by the reference's disabled
        change-propagation code, train_test_graph.py:151-157)"""
        indptr, indices, _ = self._csr_cache()
        return indices[int(indptr[v]):int(indptr[v + 1])]

    def in_degree(self, v):
        indptr, _, _ = self._csr_cache()
        return int(indptr[v + 1] - indptr[v])

    def out_degree(self, v):
        """out-degree == in-degree here: every stream of the reference is symmetrised (both directions are inserted,
        dynamic_graph_edge.py:214-215 / dgl.from_networkx of an undirected graph)"""
        return self.in_degree(v)

    def _csr_cache(self):
        key = (self.native.num_vertices, self.native.num_edges)
        if getattr(self, "_csr_key", None) != key:
            self._csr = tuple(t.cpu() for t in self.native.export_csr(with_eids=False)[:2]) + (None,)
            self._csr_key = key
        return self._csr

    def add_nodes(self, n, data=None):
        row0 = self.number_of_nodes()
        self.native.insert_vertices(n)
        if data:
            feat, target = data.get("feat"), data.get("target")
            if n > 0:
                self.features.write(row0, feat, target)
                if feat is not None and self._feat is not None:
                    self._feat[row0:row0 + n] = torch.as_tensor(feat).to("cuda", torch.float32)
                if target is not None:
                    self._target[row0:row0 + n] = torch.as_tensor(target).to("cuda", torch.int64).reshape(-1, 1)

    def add_edges(self, u, v, data=None, symmetric=False):
        self.native.insert_edges(u, v, symmetric=symmetric)

    def csr(self):
        """canonical (indptr, indices, eids) int64 CUDA tensors"""
        return self.native.export_csr()

    def sampling_plan(self, hop_fanouts, max_seeds):
        key = (tuple(hop_fanouts), int(max_seeds))
        if key not in self._splan:
            self._splan[key] = Plan([8] * (len(hop_fanouts) + 1), hop_fanouts, max_seeds, self.v_cap, mode=self.mode,
                                    seed=config.seed())
        return self._splan[key]


def edges_from(obj):
    """(src, dst) int64 numpy arrays from a DataFrame-like / dict / pair."""
    if isinstance(obj, (tuple, list)) and len(obj) == 2:
        s, d = obj
    else:
        s, d = obj["src"], obj["dst"]
    s = getattr(s, "values", s)
    d = getattr(d, "values", d)
    if isinstance(s, torch.Tensor):
        s, d = s.cpu().numpy(), d.cpu().numpy()
    return np.ascontiguousarray(s, dtype=np.int64), np.ascontiguousarray(d, dtype=np.int64)
