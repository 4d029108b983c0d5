"""Import the reference's own Python modules (build container only: /root/reference is
absent on the GPU box) behind sys.modules stand-ins for its missing third-party imports.
Used ONLY by make_golden.py to produce the committed fixtures."""
import sys
import types
import numpy as np
import torch

REF = "/root/reference/train"


class _RecGraph:
    """Recording mini-dgl graph: just enough surface for train/graph/*.py."""

    def __init__(self, src=None, dst=None, n=0):
        self.src = [] if src is None else list(src)
        self.dst = [] if dst is None else list(dst)
        self.n = n
        self.ndata, self.edata = {}, {}

    def add_nodes(self, n, data=None):
        self.n += n
        for k, v in (data or {}).items():
            self.ndata[k] = v if k not in self.ndata else torch.cat([self.ndata[k], v])

    def add_edges(self, u, v, data=None):
        self.src += [int(x) for x in u]
        self.dst += [int(x) for x in v]

    def nodes(self):
        return torch.arange(self.n)

    def number_of_edges(self):
        return len(self.src)

    def __len__(self):
        return self.n

    def subgraph(self, nodes):
        sg = _RecGraph(n=len(nodes))
        sg.ndata["_ID"] = torch.tensor(list(nodes), dtype=torch.int64)
        return sg


def load_reference():
    if "tensorflow" not in sys.modules:
        for name in ("tensorflow", "matplotlib", "matplotlib.pyplot", "seaborn"):
            sys.modules[name] = types.ModuleType(name)
        dgl = types.ModuleType("dgl")
        dgl.NID, dgl.EID = "_ID", "_EID"
        dgl.graph = lambda edges=None: _RecGraph()
        for sub in ("dgl.nn", "dgl.nn.pytorch", "dgl.nn.pytorch.conv", "dgl.nn.pytorch.conv.sageconv"):
            sys.modules[sub] = types.ModuleType(sub)
        sys.modules["dgl.nn.pytorch.conv.sageconv"].SAGEConv = object
        sys.modules["dgl"] = dgl
        sys.path.insert(0, REF)
    import utils
    utils.LIB = utils.Lib_supported.PYTORCH
    from prioritized_replay import segment_tree, replay_buffer
    from graph import dynamic_graph_edge, dynamic_graph_vertex, train_test_graph
    return dict(utils=utils, segment_tree=segment_tree, replay_buffer=replay_buffer,
                dynamic_graph_edge=dynamic_graph_edge, dynamic_graph_vertex=dynamic_graph_vertex,
                train_test_graph=train_test_graph, RecGraph=_RecGraph)
