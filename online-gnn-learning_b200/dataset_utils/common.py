"""Shared pieces of the dataset loaders: file checks, the adjlist -> directed edge list conversion and the two stream
constructors."""
import json
import os

import numpy as np

from ..graph.dynamic_graph_edge import DynamicGraphEdge
from ..graph.dynamic_graph_vertex import DynamicGraphVertex, ParentGraph


def require_files(path, files):
    missing = [f for f in files if not os.path.isfile(os.path.join(path, f))]
    if missing:
        raise FileNotFoundError("dataset files missing under %s: %s (the reference would download them here; this build has no "
                                "network access -- place the files, or generate a synthetic dataset of the same layout with "
                                "tools/make_synthetic_dataset.py)" % (path, ", ".join(missing)))


def read_adjlist_directed(path):
    """(src, dst, n_vertices) of `dgl.from_networkx(nx.read_adjlist(path, nodetype=int))` [DGL 0.5 semantics recalled]:
    the undirected graph becomes both directed edges, enumerated source by source in the graph's node order and, per
    source, in adjacency order (networkx `to_directed().edges()` order); that order is the DGL edge-id order.  Node labels
    that are not exactly 0..N-1 are relabelled in sorted order."""
    import networkx as nx
    G = nx.read_adjlist(path, nodetype=int)
    nodes = list(G.nodes())
    n = len(nodes)
    lut = None
    if sorted(nodes) != list(range(n)):
        lut = {v: i for i, v in enumerate(sorted(nodes))}
    src, dst = [], []
    for u, nbrs in G.adjacency():
        for v in nbrs:
            src.append(u if lut is None else lut[u])
            dst.append(v if lut is None else lut[v])
    return np.asarray(src, dtype=np.int64), np.asarray(dst, dtype=np.int64), n


def read_timestamps(path):
    with open(path) as f:
        return json.load(f, object_hook=lambda d: {int(k): v for k, v in d.items()})


def labels_and_classes(targets):
    """labelled vertices = label != -1; the class count includes the -1 "unknown" label when present, exactly like the
    reference's `len(np.unique(targets))` (elliptic reports 3 classes)"""
    targets = np.asarray(targets).astype(np.int64)
    if targets.ndim == 1:
        targets = targets.reshape(-1, 1)
    labelled = set(np.argwhere(targets != -1)[:, 0].tolist())
    return targets, labelled, len(np.unique(targets))


def vertex_stream(path, feat_file, ts_file, snapshots):
    """pubmed / elliptic / arxiv: full graph + vertex arrival times -> two DynamicGraphVertex (train and delta-ahead test)"""
    require_files(path, [feat_file, "targets.npy", "graph.adjlist", ts_file])
    feats = np.load(os.path.join(path, feat_file)).astype(np.float32)          # utils.to_nn_lib: float64 -> float32
    targets, labelled, n_classes = labels_and_classes(np.load(os.path.join(path, "targets.npy")))
    src, dst, n = read_adjlist_directed(os.path.join(path, "graph.adjlist"))
    if n != feats.shape[0]:
        raise ValueError("graph.adjlist has %d vertices but %s has %d rows" % (n, feat_file, feats.shape[0]))
    ts = read_timestamps(os.path.join(path, ts_file))
    out = []
    for _ in range(2):
        pg = ParentGraph(src, dst, n)
        pg.ndata["feat"], pg.ndata["target"] = feats, targets
        dyn = DynamicGraphVertex(pg, snapshots, labelled)
        dyn.build(vertex_timestamps=ts)
        out.append(dyn)
    return feats.shape[1], targets, out[0], n_classes, out[1]


def edge_stream(path, snapshots):
    """reddit: time-ordered edge list (columns src, dst; vertex ids dense in first-appearance order) -> two DynamicGraphEdge"""
    import pandas as pd
    require_files(path, ["feat_data.npy", "targets.npy", "edges_dataframe.csv"])
    feats = np.load(os.path.join(path, "feat_data.npy")).astype(np.float32)
    targets, labelled, n_classes = labels_and_classes(np.load(os.path.join(path, "targets.npy")))
    edges = pd.read_csv(os.path.join(path, "edges_dataframe.csv"), na_filter=False, dtype=np.int64)
    out = []
    for _ in range(2):
        dyn = DynamicGraphEdge(snapshots, labelled)
        dyn.build(feats, targets, True, edge_timestamps=edges, keep_master=False)
        out.append(dyn)
    return feats.shape[1], targets, out[0], n_classes, out[1]
