"""Prioritized rehearsal buffer over the GPU sum tree (mirror of
train/prioritized_replay/replay_buffer.py:60-282).

Quirks kept on purpose: the log-space running min/max is never reset and old
leaves are never re-normalised; eps is 1e-5 on add and 1e-6 on update; `p_total` excludes the last
leaf; leaf = v ** alpha with alpha fixed at construction.

Two ways in:
  * add_all / update_priorities(dict)   host dict API; the transform runs on the host in Python
    floats exactly like the reference (bit-identical leaves), only the tree lives on the GPU;
  * update_from_losses(ids, loss_tensor) device API for the PBR trainer: clip/log/normalise/pow run in
    one kernel and the per-vertex losses never leave the GPU (libm-level rounding differences only).
"""
import math
import random

import numpy as np
import torch

from .segment_tree import SumSegmentTree


class ReplayBuffer:
    def __init__(self, size):
        self._storage = []
        self._maxsize = size
        self._next_idx = 0

    def __len__(self):
        return len(self._storage)

    def add(self, node_id):
        self._storage.append(node_id)
        self._next_idx += 1

    def _encode_sample(self, idxes):
        return np.array([self._storage[i] for i in idxes])

    def sample(self, batch_size):
        idxes = [random.randint(0, len(self._storage) - 1) for _ in range(batch_size)]
        return self._encode_sample(idxes)


class PrioritizedReplayBuffer(ReplayBuffer):
    def __init__(self, size, alpha, max_priority, min_priority, tree_backend=None, verbose=True):
        super().__init__(size)
        if alpha < 0:
            raise AssertionError("alpha must be >= 0")
        self._alpha = alpha
        cap = 1
        while cap < size:
            cap *= 2
        self._it_sum = SumSegmentTree(cap, backend=tree_backend(cap) if tree_backend else None)
        self._max_clip_priority = max_priority
        self._min_clip_priority = min_priority
        self._key_to_idx = {}
        self._max_priority, self._min_priority = -1, 99999999
        self.max_val, self.min_val = -1, 99999999
        self._dev_state = None
        if verbose:
            print("PrioritizedBuffer init, alpha: ", self._alpha)

    def get_max_priority(self):
        self._pull_dev_state()
        return self.max_val

    def get_min_priority(self):
        self._pull_dev_state()
        return self.min_val

    def _pull_dev_state(self):
        """device-side updates (update_from_losses) advance the running min / max on the GPU only: bring them back (one 4-double
        D2H copy) before any host-side reader or writer of that state runs, then let the host copy be the truth again"""
        if self._dev_state is not None:
            self.sync_state()
            self._dev_state = None

    def _check_room(self, n_new):
        cap = self._it_sum.capacity
        if self._next_idx + n_new > cap:
            raise AssertionError("PrioritizedReplayBuffer: %d + %d entries exceed the tree capacity %d (the reference asserts "
                                 "0 <= idx < capacity in SegmentTree.__setitem__)" % (self._next_idx, n_new, cap))

    # ---- host (dict) API: transform in Python floats, tree on the device -----------------------
    def _normalize(self, node_priority_dict):
        logs = {}
        lo, hi = self._min_clip_priority, self._max_clip_priority
        for node, p in node_priority_dict.items():
            p = min(max(p, lo), hi)
            if p > self.max_val:
                self.max_val = p
            if p < self.min_val:
                self.min_val = p
            lp = math.log(p)
            logs[node] = lp
            if lp > self._max_priority:
                self._max_priority = lp
            if lp < self._min_priority:
                self._min_priority = lp
        return logs

    def _leaf_values(self, logs, eps):
        scale = self._max_priority - self._min_priority
        out = []
        for lp in logs.values():
            v = (lp - self._min_priority) / scale if scale > 0 else (lp - self._min_priority)
            v += eps
            if v < 0:
                raise AssertionError("negative normalised priority")
            out.append(v ** self._alpha)
        return out

    def add_all(self, node_priority_dict):
        self._pull_dev_state()
        self._check_room(len(node_priority_dict))
        logs = self._normalize(node_priority_dict)
        vals = self._leaf_values(logs, 0.00001)
        idx = []
        for node in logs:
            i = self._next_idx
            self.add(node)
            self._key_to_idx[node] = i
            idx.append(i)
        if idx:
            self._it_sum.set_many(idx, vals)

    def update_priorities(self, d_priorities):
        self._pull_dev_state()
        logs = self._normalize(d_priorities)
        vals = self._leaf_values(logs, 0.000001)
        idx = [self._key_to_idx[node] for node in logs]
        if idx:
            self._it_sum.set_many(idx, vals)

    # ---- device API ---------------------------------------------------------------------------------
    def update_from_losses(self, nodes, losses_dev, adding=False):
        """nodes: iterable of vertex ids (host), losses_dev: CUDA fp32 tensor aligned with it."""
        if self._dev_state is None:
            self._dev_state = torch.tensor([self.min_val, self.max_val, self._min_priority, self._max_priority],
                                           dtype=torch.float64, device="cuda")
        if adding:
            nodes = list(nodes)
            self._check_room(len(nodes))
            idx = []
            for node in nodes:
                self._key_to_idx[node] = self._next_idx
                idx.append(self._next_idx)
                self.add(node)
        else:
            idx = [self._key_to_idx[n] for n in nodes]
        self._it_sum._t.set_from_loss(idx, losses_dev, float(self._min_clip_priority), float(self._max_clip_priority),
                                      0.00001 if adding else 0.000001, float(self._alpha), self._dev_state)

    def sync_state(self):
        """pull the running min/max back after device-side updates"""
        if self._dev_state is not None:
            s = self._dev_state.tolist()
            self.min_val, self.max_val, self._min_priority, self._max_priority = s

    # ---- sampling -------------------------------------------------------------------------------------
    def _sample_proportional(self, batch_size):
        """Stratified proportional draw (reference :164-203); consumes the Python `random` stream in the same
        order (batch_size uniforms, then the top-up / fill draws) so it is reproducible under the same seed."""
        n_items = len(self._storage)
        if batch_size >= n_items:
            return list(self._key_to_idx.values())
        u = [random.random() for _ in range(batch_size)]
        found = self._it_sum._t.sample_stratified(u, n_items).tolist()
        res = set(found)
        if len(res) < batch_size:
            p_total = self._it_sum.sum(0, n_items - 1)
            tries = 0
            while len(res) < batch_size:
                res.add(self._it_sum.find_prefixsum_idx(random.random() * p_total))
                tries += 1
                if tries > 20:
                    break
            while len(res) < batch_size:
                res.add(random.randint(0, n_items - 1))
        return res

    def sample(self, batch_size):
        return list(self._encode_sample(self._sample_proportional(batch_size)))

    def increment_priorities(self, node, increment):
        if increment < 0:
            raise AssertionError("increment must be >= 0")
        self._pull_dev_state()
        idx = self._key_to_idx[node]
        cur = self._it_sum[idx]
        if self._max_priority == -1:
            cur += increment ** self._alpha
        else:
            cur += increment * (self._max_priority - self._min_priority)
        self._it_sum[idx] = min(cur, 1)

    def dump_priorities(self, vertex_list):
        idx = [self._key_to_idx[v] for v in vertex_list]
        return self._it_sum.get_many(idx).tolist() if idx else []
