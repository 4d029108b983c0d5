"""reddit loader (mirror of train/dataset_utils/reddit.py:144-177): feat_data.npy, targets.npy, edges_dataframe.csv -> edge
stream (the edge list must already be relabelled to first-appearance order, reddit.py:87-141 `relabel`)."""
from .common import edge_stream

FILES = ["feat_data.npy", "targets.npy", "edges_dataframe.csv"]


def load(path, snapshots=100, cuda=True, copy_to_gpu=True):
    return edge_stream(path, snapshots)
