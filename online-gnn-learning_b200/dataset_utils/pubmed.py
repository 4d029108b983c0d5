"""pubmed loader (mirror of train/dataset_utils/pubmed.py:70-124): feat_data.npy, targets.npy, graph.adjlist, postponed_timestamp.json -> vertex stream."""
from .common import vertex_stream

FILES = ["feat_data.npy", "targets.npy", "graph.adjlist", "postponed_timestamp.json"]


def load(path, snapshots=100, cuda=True, copy_to_gpu=True):
    return vertex_stream(path, "feat_data.npy", "postponed_timestamp.json", snapshots)
