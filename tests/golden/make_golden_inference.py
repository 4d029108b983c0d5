"""Golden fixture of the cached streaming-inference path, produced by RUNNING THE REFERENCE'S OWN HANDLER CODE
(/root/reference/inference_optimized.py, `MNISTDigitClassifier.inference`, lines 144-301) in the build container.

    python tests/golden/make_golden_inference.py      ->  tests/golden/inference_stream.npz

The handler is a TorchServe plug-in with hard-coded cluster paths; only its per-request method is exercised.  Stand-ins:
  * `torchvision`, `tensorflow`, ... : empty modules (imported at module top, never used by `inference`)
  * `dgl`: a mini multigraph (`_ServeGraph`) with exactly the calls the method makes -- add_nodes (zero-filling the fields it is
    not given, as DGL does), add_edge, out_degrees, out_edges, in_edges, subgraph (induced, parallel edges kept, parent edge
    order), update_all(message UDF, reduce UDF) with DGL's degree bucketing (mailbox [nodes, degree, F], zero rows for
    in-degree 0)
  * `open()` of the handler's result file: an in-memory buffer
The object is built without `initialize()` (which loads a dataset and a checkpoint from absolute paths): its fields are set
directly to random fc_pool / fc_self / fc_neigh `torch.nn.Linear` layers and a random feature table.
"""
import builtins
import contextlib
import importlib.util
import io
import json
import os
import sys
import types

import numpy as np
import torch

OUT = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"


class _Frame:
    def __init__(self, data):
        self._d = data

    def __getitem__(self, k):
        return self._d[k]


class _ServeGraph:
    """multigraph with insertion-ordered edges and per-node tensors"""

    def __init__(self):
        self.n = 0
        self.src, self.dst = [], []
        self.ndata = {}
        self.subgraph_calls = []

    def __len__(self):
        return self.n

    def add_nodes(self, n, data=None):
        data = data or {}
        for k, v in self.ndata.items():                     # fields not supplied are zero-filled for the new nodes
            if k not in data:
                self.ndata[k] = torch.cat([v, torch.zeros((n,) + tuple(v.shape[1:]), dtype=v.dtype)])
        for k, v in data.items():
            self.ndata[k] = v if k not in self.ndata else torch.cat([self.ndata[k], v])
        self.n += n

    def add_edge(self, u, v):
        self.src.append(int(u))
        self.dst.append(int(v))

    def out_degrees(self, v):
        deg = np.bincount(np.asarray(self.src, dtype=np.int64), minlength=self.n)
        return torch.from_numpy(deg[np.asarray(v, dtype=np.int64)])

    def out_edges(self, v):
        s, d = [], []
        for x in [int(t) for t in v]:
            for a, b in zip(self.src, self.dst):
                if a == x:
                    s.append(a)
                    d.append(b)
        return torch.tensor(s, dtype=torch.int64), torch.tensor(d, dtype=torch.int64)

    def in_edges(self, v):
        s, d = [], []
        for x in [int(t) for t in v]:
            for a, b in zip(self.src, self.dst):
                if b == x:
                    s.append(a)
                    d.append(b)
        return torch.tensor(s, dtype=torch.int64), torch.tensor(d, dtype=torch.int64)

    def subgraph(self, nodes):
        nodes = [int(x) for x in nodes]
        self.subgraph_calls.append(list(nodes))
        return _SubGraph(self, nodes)


class _SubGraph:
    def __init__(self, parent, nodes):
        self.nodes = nodes
        local = {v: i for i, v in enumerate(nodes)}
        self.es = [local[a] for a, b in zip(parent.src, parent.dst) if a in local and b in local]
        self.ed = [local[b] for a, b in zip(parent.src, parent.dst) if a in local and b in local]
        idx = torch.tensor(nodes, dtype=torch.int64)
        self.ndata = {k: v[idx].clone() for k, v in parent.ndata.items()}
        self.ndata["_ID"] = idx
        self.dstdata = self.ndata

    def to(self, device):
        return self

    def update_all(self, msg, red):
        n = len(self.nodes)
        es = torch.tensor(self.es, dtype=torch.int64)
        m = msg(types.SimpleNamespace(src=_Frame({k: v[es] for k, v in self.ndata.items() if k != "_ID"})))["m"] if len(es) else None
        per_dst = [[] for _ in range(n)]
        for e, d in enumerate(self.ed):
            per_dst[d].append(e)
        out = None
        by_deg = {}
        for v, lst in enumerate(per_dst):
            if lst:
                by_deg.setdefault(len(lst), []).append(v)
        results = {}
        for deg, vs in sorted(by_deg.items()):
            mailbox = torch.stack([m[torch.tensor(per_dst[v])] for v in vs])          # [nodes, degree, F]
            r = red(types.SimpleNamespace(mailbox={"m": mailbox}))
            for k, val in r.items():
                results.setdefault(k, []).append((vs, val))
        for k, parts in results.items():
            width = parts[0][1].shape[1]
            full = torch.zeros(n, width)
            for vs, val in parts:
                full[torch.tensor(vs)] = val
            self.ndata[k] = full
        if not results:                                       # no edge inside the induced subgraph: zero field (name from the UDF)
            probe = red(types.SimpleNamespace(mailbox={"m": torch.zeros(1, 1, self._width_hint)}))
            for k in probe:
                self.ndata[k] = torch.zeros(n, self._width_hint)


def load_handler():
    for name in ("torchvision", "torchvision.transforms", "tensorflow", "matplotlib", "matplotlib.pyplot", "seaborn", "dgl.nn", "dgl.nn.pytorch",
                 "dgl.nn.pytorch.conv", "dgl.nn.pytorch.conv.sageconv"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["dgl.nn.pytorch.conv.sageconv"].SAGEConv = object
    sys.modules["torchvision"].transforms = sys.modules["torchvision.transforms"]
    dgl = types.ModuleType("dgl")
    dgl.NID, dgl.EID = "_ID", "_EID"
    dgl.DGLGraph = _ServeGraph
    sys.modules["dgl"] = dgl
    sys.path.insert(0, os.path.join(REF_ROOT, "train"))
    real_open = builtins.open

    def fake_open(path, *a, **k):
        if isinstance(path, str) and path.startswith("/project/"):
            return io.StringIO()
        return real_open(path, *a, **k)

    spec = importlib.util.spec_from_file_location("ref_inference_optimized", os.path.join(REF_ROOT, "inference_optimized.py"))
    mod = importlib.util.module_from_spec(spec)
    builtins.open = fake_open
    try:
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            spec.loader.exec_module(mod)
    finally:
        builtins.open = real_open
    return mod


def request_stream(rng, n_vertices, n_requests):
    """edge batches over vertex ids that grow with time (the handler sizes the graph by max id + 1); both directions occur so
    that successor sets are non-empty; one hub crosses the handler's out-degree threshold of 15"""
    reqs = []
    hi = 6
    hub = 2
    for r in range(n_requests):
        hi = min(n_vertices, hi + int(rng.integers(0, 4)))
        k = int(rng.integers(1, 5))
        edges = []
        for _ in range(k):
            a, b = int(rng.integers(0, hi)), int(rng.integers(0, hi))
            edges.append([a, b])
            if rng.random() < 0.6:
                edges.append([b, a])
        if r % 3 == 0:                                       # feed the hub's out-degree (edges are stored reversed: [dst, src])
            edges.append([int(rng.integers(0, hi)), hub])
            edges.append([hub, int(rng.integers(0, hi))])
        reqs.append(edges)
    return reqs


def main():
    mod = load_handler()
    torch.manual_seed(7)
    rng = np.random.default_rng(7)
    V, F, H, C = 48, 12, 8, 4
    feat = torch.randn(V, F)
    target = torch.randint(0, C, (V, 1))
    dims = [(F, H), (H, C)]
    fc_pool = [torch.nn.Linear(F, F), torch.nn.Linear(H, H)]
    fc_self = [torch.nn.Linear(i, o) for i, o in dims]
    fc_neigh = [torch.nn.Linear(i, o) for i, o in dims]
    h = object.__new__(mod.MNISTDigitClassifier)
    h.fc_pool, h.fc_self, h.fc_neigh = fc_pool, fc_self, fc_neigh
    h.graph = _ServeGraph()
    h.graph_feat = types.SimpleNamespace(ndata={"feat": feat.clone(), "target": target.clone()})   # DGL node data are tensors
    h.cuda = False
    h.copy_dataset_gpu = False
    h.requests = 0
    h.file_results = io.StringIO()
    _SubGraph._width_hint = 1
    reqs = request_stream(rng, V, 40)
    rec = dict(edges=[], n_edges=[], P=[], nP=[], S=[], nS=[], out=[], caches={k: [] for k in ("h0proj", "neigh0", "h1", "h1proj", "neigh1", "h2")},
               n_nodes=[])
    with torch.no_grad():
        for edges in reqs:
            h.graph.subgraph_calls.clear()
            with contextlib.redirect_stdout(io.StringIO()):
                ans = h.inference([{"body": json.dumps(edges)}])
            out = json.loads(ans[0])
            P, S = h.graph.subgraph_calls[0], h.graph.subgraph_calls[1]
            assert len(out) == len(P)
            rec["edges"] += [e for e in edges]
            rec["n_edges"].append(len(edges))
            rec["P"] += P
            rec["nP"].append(len(P))
            rec["S"] += S
            rec["nS"].append(len(S))
            rec["out"] += out
            rec["n_nodes"].append(len(h.graph))
            for k in rec["caches"]:
                full = torch.zeros(V, h.graph.ndata[k].shape[1])
                full[:len(h.graph)] = h.graph.ndata[k]
                rec["caches"][k].append(full.numpy().copy())
    params = {}
    for l in range(2):
        for name, layer in (("fc_pool", fc_pool[l]), ("fc_self", fc_self[l]), ("fc_neigh", fc_neigh[l])):
            params["layers.%d.%s.weight" % (l, name)] = layer.weight.detach().numpy().copy()
            params["layers.%d.%s.bias" % (l, name)] = layer.bias.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "inference_stream.npz"), feat=feat.numpy(), edges=np.array(rec["edges"], dtype=np.int64),
                        n_edges=np.array(rec["n_edges"]), P=np.array(rec["P"], dtype=np.int64), nP=np.array(rec["nP"]),
                        S=np.array(rec["S"], dtype=np.int64), nS=np.array(rec["nS"]), out=np.array(rec["out"], dtype=np.int64),
                        n_nodes=np.array(rec["n_nodes"]), **{"cache_" + k: np.stack(v) for k, v in rec["caches"].items()},
                        **{"param_" + k: v for k, v in params.items()})
    print("requests", len(reqs), "edges", len(rec["edges"]), "mean |P|", np.mean(rec["nP"]), "mean |S|", np.mean(rec["nS"]),
          "final nodes", rec["n_nodes"][-1])


if __name__ == "__main__":
    main()
