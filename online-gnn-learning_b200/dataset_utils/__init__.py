"""Loaders of the reference's on-disk dataset formats: each `load(path, snapshots, cuda, copy_to_gpu)`
returns `(feat_size, targets, dynamic_graph, n_classes, dynamic_graph_test)` like train/dataset_utils/*.py:load, with the
graph state living in the device-resident streaming CSR instead of DGL objects.  The reference's downloaders and
preprocessors (raw dumps -> these files) are not rebuilt: there is no network here and they are offline tools."""
from . import common, pubmed, bitcoin, arxiv, reddit     # noqa: F401

LOADERS = {"pubmed": pubmed.load, "elliptic": bitcoin.load, "arxiv": arxiv.load, "reddit": reddit.load}
