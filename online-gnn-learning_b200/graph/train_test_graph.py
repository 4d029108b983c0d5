"""Train/test split + rehearsal vertex choosers over a DynamicGraph (mirror of
train/graph/train_test_graph.py:12-248).

faithful mode (config.set_faithful(True), default) reproduces the reference's draws literally -- numpy
train_test_split, Python-set iteration order, in-place random.shuffle, and the PBR quirk that
draw_priority_train_nodes is a uniform shuffle whenever n <= |train| -- so the vertex
sequences match the reference under random.seed / np.random.seed.  fast mode keeps the train set as a
device tensor and draws with the counter-RNG kernels (uniform subset / real stratified proportional).
"""
import random
from itertools import compress

import numpy as np
import torch
from sklearn.model_selection import train_test_split

from .. import config
from .._native import draw_uniform
from ..prioritized_replay.replay_buffer import PrioritizedReplayBuffer

SIZE_BUFFER = 10000000


class TrainTestGraph:
    def __init__(self, graph, split=0.25, start_prior_alpha=1, end_prior_alpha=2, scale=1, max_priority=3.0,
                 start_priority=2, min_priority=0.0000001, tree_backend=None, verbose=False):
        self.scale = scale
        self.temporal_graph = graph
        self.train_set, self.test_set = set(), set()
        self.size_evolution = len(graph)
        self.split = split
        self.graph = graph.get_graph()
        self.prior_alpha = start_prior_alpha
        self.start_prior_alpha, self.end_prior_alpha = start_prior_alpha, end_prior_alpha
        self.max_priority, self.start_priority, self.min_priority = max_priority, start_priority, min_priority
        self._tree_backend, self._verbose = tree_backend, verbose
        self._draw_counter = 0
        self._train_dev = None
        self.priority_replay_buffer = self._new_buffer()
        added, labelled = graph.get_added_vertices()
        self._draw_train_test(list(compress(added, labelled)))

    def _new_buffer(self):
        return PrioritizedReplayBuffer(SIZE_BUFFER, self.prior_alpha, max_priority=self.max_priority,
                                       min_priority=self.min_priority, tree_backend=self._tree_backend, verbose=self._verbose)

    # ---- split --------------------------------------------------------------------------------------
    def _draw_train_test(self, vertices):
        if len(vertices) >= 3:
            self.train, self.test = train_test_split(vertices, shuffle=True, test_size=self.split)
        else:
            self.train, self.test = set(vertices), set()
        self.train_set = self.train_set.union(set(self.train))
        self.train_set_list = list(self.train_set)
        self.test_set = self.test_set.union(set(self.test))
        self.test_set_list = list(self.test_set)
        self._train_dev = None
        self._update_priority_struct()

    def _update_priority_struct(self):
        buf = self.priority_replay_buffer
        lo, hi = buf.get_min_priority(), buf.get_max_priority()
        first = hi == -1
        p_new = self.start_priority if first else lo + (hi - lo) * 0.95
        buf.add_all({v: p_new for v in self.train})

    def __len__(self):
        return len(self.temporal_graph)

    def _get_affected_nodes(self, source_node, depth=2):
        """{vertex: priority bump} of the vertices within `depth` in-hops of a changed vertex, each hop scaled by
        1 / out_degree and `scale`, capped at 1 (mirror of train_test_graph.py:139-166; the reference's only caller is
        commented out, so this is offered for completeness -- SURVEY 8(f)-4)."""
        g = self.temporal_graph.get_graph()
        nbrs = {source_node: 1 * self.scale}
        for _ in range(depth):
            # ONE neighbourhood query per hop for the whole frontier (ogl_graph_gather_rows: the in-edge rows of all frontier
            # vertices, each in edge-id order) + one degree query for the distinct predecessors; the float accumulation below then
            # runs in the reference's dict order
            keys = list(nbrs.keys())
            offsets, src = g.native.gather_rows(np.asarray(keys, dtype=np.int64))
            offsets, src = offsets.cpu().numpy(), src.cpu().numpy()
            uniq = np.unique(src)
            # out-degree == in-degree: every stream of the reference is symmetrised (DeviceGraph.out_degree)
            out_deg = dict(zip(uniq.tolist(), g.native.row_degrees(uniq).cpu().tolist())) if len(uniq) else {}
            tmp = {}
            for i, k in enumerate(keys):
                v = nbrs[k]
                for nbr in src[offsets[i]:offsets[i + 1]].tolist():
                    w = (1 / out_deg[nbr]) * v * self.scale
                    if nbr in tmp:
                        tmp[nbr] = min(tmp[nbr] + w, 1)
                    else:
                        tmp[nbr] = w
            for k, v in tmp.items():
                nbrs[k] = max(nbrs[k], v) if k in nbrs else v
        return nbrs

    def evolve(self):
        g = self.temporal_graph
        span = self.end_prior_alpha - self.start_prior_alpha
        self.prior_alpha = self.start_prior_alpha + (span / len(self)) * g.evolution_index
        g.evolve()
        self.graph = g.get_graph()
        added, labelled = g.get_added_vertices()
        self._draw_train_test(list(compress(added, labelled)))

    # ---- getters ------------------------------------------------------------------------------------
    def get_graph(self):
        return self.temporal_graph.get_graph()

    def get_train_set(self):
        return self.train_set_list

    def get_test_set(self):
        return self.test_set_list

    def get_new_test_nodes(self):
        return self.test

    def get_new_train_nodes(self, batch_size=None):
        fresh = list(self.train)
        if batch_size is None or batch_size >= len(fresh):
            return fresh
        random.shuffle(fresh)
        return fresh[:batch_size]

    def get_original_to_subgraph_map(self):
        return self.temporal_graph.get_original_to_subgraph_map()

    def get_subgraph_to_original_map(self):
        return self.temporal_graph.get_subgraph_to_original_map()

    # ---- choosers -----------------------------------------------------------------------------------
    def _train_tensor(self):
        if self._train_dev is None:
            self._train_dev = torch.as_tensor(np.asarray(self.train_set_list, dtype=np.int64), device="cuda")
        return self._train_dev

    def draw_random_train_nodes(self, n_nodes):
        """RBR: uniform n-subset of the train set (whole list if it is smaller)."""
        if n_nodes > len(self.train_set_list):
            return self.train_set_list
        if config.faithful():
            random.shuffle(self.train_set_list)
            self._train_dev = None
            return self.train_set_list[:n_nodes]
        pop = self._train_tensor()
        self._draw_counter += 1
        return pop[draw_uniform(pop.numel(), n_nodes, config.seed(), self._draw_counter)]

    def draw_priority_train_nodes(self, n_nodes):
        """PBR draw.  faithful: identical to the RBR shuffle when n <= |train| (reference :218-223)."""
        if config.faithful():
            if n_nodes <= len(self.train_set_list):
                random.shuffle(self.train_set_list)
                self._train_dev = None
                return self.train_set_list[:n_nodes]
            return self.priority_replay_buffer.sample(n_nodes)
        return self.priority_replay_buffer.sample(n_nodes)

    def dump_priorities(self, vertex_list):
        return self.priority_replay_buffer.dump_priorities(vertex_list)

    def update_priorities(self, d_priorities):
        """partial update -> tree update; an update covering the whole train set rebuilds the buffer with the
        current annealed alpha (reference :228-242)."""
        if len(d_priorities) > len(self.train_set):
            raise AssertionError("more priorities than train vertices")
        if len(d_priorities) < len(self.train_set):
            self.priority_replay_buffer.update_priorities(d_priorities)
        else:
            self.priority_replay_buffer = self._new_buffer()
            self.priority_replay_buffer.add_all(d_priorities)

    def update_priorities_device(self, nodes, losses_dev):
        """device twin of update_priorities: losses stay on the GPU"""
        if len(nodes) < len(self.train_set):
            self.priority_replay_buffer.update_from_losses(nodes, losses_dev, adding=False)
        else:
            self.priority_replay_buffer = self._new_buffer()
            self.priority_replay_buffer.update_from_losses(nodes, losses_dev, adding=True)
