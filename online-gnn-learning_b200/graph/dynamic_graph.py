"""Streaming-graph interface (mirror of the reference's train/graph/dynamic_graph.py:7-43)."""


class DynamicGraph:
    """A graph that arrives as `snapshots` slices.  Label -1 == unlabelled."""

    def __init__(self, graph, snapshots, labelled_vertices, search_depth):
        if snapshots <= 0:
            raise AssertionError("snapshots must be >= 1")
        self.graph = graph
        self.snapshots = snapshots
        self.search_depth = search_depth
        self.evolution_index = 0
        self.labelled_vertices = labelled_vertices

    def get_labelled_vertices(self):
        return self.labelled_vertices

    def get_added_vertices(self, delta=None):
        raise NotImplementedError

    def get_graph(self):
        raise NotImplementedError

    def __len__(self):
        raise NotImplementedError

    def evolve(self):
        raise NotImplementedError


class IdentityMap:
    """subgraph id == original id (edge streams; reference `Wrap`, dynamic_graph_edge.py:263-265)."""

    def __getitem__(self, item):
        return item


def labelled_mask(labelled_vertices, vertices):
    """[v in labelled_vertices for v in vertices] for a set or a boolean/0-1 array."""
    import numpy as np
    if isinstance(labelled_vertices, (set, frozenset, dict)):
        return [v in labelled_vertices for v in vertices]
    arr = np.asarray(labelled_vertices)
    v = np.asarray(vertices, dtype=np.int64)
    if arr.dtype == np.bool_:
        return arr[v].tolist()
    return np.isin(v, arr).tolist()
