"""Sum-tree + prioritized buffer -- oracle only (numpy fp64, IEEE-exact restatement).

Follows train/prioritized_replay/segment_tree.py:4-125 and replay_buffer.py:60-244.
PINNED: tests/test_oracle_pinning.py runs this against the reference's own classes
(fixtures in tests/golden/replay_*.npz were produced by the reference code itself).
"""
import math
import numpy as np


class SumTree:
    """Array tree value[2*cap]; leaf i lives at cap+i; parent = left + right
    (segment_tree.py:69-79)."""

    def __init__(self, capacity):
        assert capacity > 0 and capacity & (capacity - 1) == 0
        self.cap = capacity
        self.value = np.zeros(2 * capacity, dtype=np.float64)

    def set(self, idx, val):
        """Batch leaf write; later duplicates win (sequential __setitem__ semantics)."""
        idx = np.asarray(idx, dtype=np.int64)
        val = np.asarray(val, dtype=np.float64)
        if len(idx) == 0:
            return
        last = {}
        for i, v in zip(idx.tolist(), val.tolist()):
            last[i] = v
        ii = np.fromiter(last.keys(), dtype=np.int64) + self.cap
        self.value[ii] = np.fromiter(last.values(), dtype=np.float64)
        node = np.unique(ii >> 1)
        while node.size and node[0] >= 1:
            self.value[node] = self.value[2 * node] + self.value[2 * node + 1]
            node = np.unique(node >> 1)
            if node[0] == 0:
                node = node[1:]

    def _reduce(self, start, end, node, ns, ne):          # segment_tree.py:30-45 (end inclusive)
        if start == ns and end == ne:
            return float(self.value[node])
        mid = (ns + ne) // 2
        if end <= mid:
            return self._reduce(start, end, 2 * node, ns, mid)
        if mid + 1 <= start:
            return self._reduce(start, end, 2 * node + 1, mid + 1, ne)
        return self._reduce(start, mid, 2 * node, ns, mid) + self._reduce(mid + 1, end, 2 * node + 1, mid + 1, ne)

    def sum(self, start=0, end=None):                      # [start, end)  (segment_tree.py:47-67)
        if end is None:
            end = self.cap
        if end < 0:
            end += self.cap
        end -= 1
        return self._reduce(start, end, 1, 0, self.cap - 1)

    def find_prefixsum_idx(self, mass):                    # segment_tree.py:118-125
        mass = np.array(mass, dtype=np.float64, ndmin=1).copy()
        idx = np.ones(len(mass), dtype=np.int64)
        while idx[0] < self.cap:
            left = self.value[2 * idx]
            go_left = left > mass
            mass = np.where(go_left, mass, mass - left)
            idx = np.where(go_left, 2 * idx, 2 * idx + 1)
        return idx - self.cap


class PrioritizedBufferOracle:
    """PrioritizedReplayBuffer restated (replay_buffer.py:60-244); same quirks: running
    log-space min/max never reset, eps 1e-5 on add vs 1e-6 on update, leaf = v**alpha."""

    def __init__(self, size, alpha, max_priority, min_priority):
        cap = 1
        while cap < size:
            cap *= 2
        self.tree = SumTree(cap)
        self.alpha = alpha
        self.max_clip, self.min_clip = max_priority, min_priority
        self.storage = []
        self.key_to_idx = {}
        self._max_priority, self._min_priority = -1, 99999999
        self.max_val, self.min_val = -1, 99999999

    def __len__(self):
        return len(self.storage)

    def _normalize(self, d):                               # :110-130
        out = {}
        for node, p in d.items():
            p = min(max(p, self.min_clip), self.max_clip)
            self.max_val = max(self.max_val, p) if p > self.max_val else self.max_val
            self.min_val = p if p < self.min_val else self.min_val
            l = math.log(p)
            out[node] = l
            if l > self._max_priority:
                self._max_priority = l
            if l < self._min_priority:
                self._min_priority = l
        return out

    def _leaf(self, l, eps):
        scale = self._max_priority - self._min_priority
        v = (l - self._min_priority) / scale if scale > 0 else (l - self._min_priority)
        v += eps
        return v ** self.alpha

    def add_all(self, d):                                  # :133-160
        dn = self._normalize(d)
        idx, val = [], []
        for node, l in dn.items():
            i = len(self.storage)
            self.storage.append(node)
            self.key_to_idx[node] = i
            idx.append(i)
            val.append(self._leaf(l, 0.00001))
        self.tree.set(idx, val)

    def update_priorities(self, d):                        # :219-244
        dn = self._normalize(d)
        idx = [self.key_to_idx[n] for n in dn]
        val = [self._leaf(l, 0.000001) for l in dn.values()]
        self.tree.set(idx, val)

    def sample_proportional(self, n, uniforms, topup_uniforms=(), fill_randints=()):
        """_sample_proportional (:164-203) with the `random` draws supplied by the caller:
        uniforms[n] for the stratified pass, then up to 21 top-up uniforms, then randint
        values.  Returns the list of tree indices in first-insertion order (the reference
        returns a set; callers compare as sets)."""
        if n >= len(self.storage):
            return list(self.key_to_idx.values())
        p_total = self.tree.sum(0, len(self.storage) - 1)  # excludes the last leaf (:169)
        every = p_total / n
        res = []
        seen = set()
        for i in range(n):
            mass = uniforms[i] * every + i * every
            idx = int(self.tree.find_prefixsum_idx(mass)[0])
            if idx not in seen:
                seen.add(idx); res.append(idx)
        j = 0
        it = iter(topup_uniforms)
        while len(seen) < n:
            mass = next(it) * p_total
            idx = int(self.tree.find_prefixsum_idx(mass)[0])
            if idx not in seen:
                seen.add(idx); res.append(idx)
            j += 1
            if j > 20:
                break
        it = iter(fill_randints)
        while len(seen) < n:
            idx = int(next(it))
            if idx not in seen:
                seen.add(idx); res.append(idx)
        return res
