"""arxiv loader (mirror of train/dataset_utils/arxiv.py:18-54): feats.npy, targets.npy, graph.adjlist, vertex_timestamp.json -> vertex stream."""
from .common import vertex_stream

FILES = ["feats.npy", "targets.npy", "graph.adjlist", "vertex_timestamp.json"]


def load(path, snapshots=100, cuda=True, copy_to_gpu=True):
    return vertex_stream(path, "feats.npy", "vertex_timestamp.json", snapshots)
