"""bitcoin loader (mirror of train/dataset_utils/bitcoin.py:78-114): feat_data.npy, targets.npy, graph.adjlist, vertex_timestamp.json -> vertex stream."""
from .common import vertex_stream

FILES = ["feat_data.npy", "targets.npy", "graph.adjlist", "vertex_timestamp.json"]


def load(path, snapshots=100, cuda=True, copy_to_gpu=True):
    return vertex_stream(path, "feat_data.npy", "vertex_timestamp.json", snapshots)
