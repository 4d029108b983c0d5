"""Sum tree on the GPU (mirror of the SumSegmentTree surface of
train/prioritized_replay/segment_tree.py:86-125; arithmetic in csrc/replay.cu, fp64, bit-exact)."""
import torch

from .._native import SumTree as _NativeTree


class SumSegmentTree:
    def __init__(self, capacity, backend=None):
        if capacity <= 0 or capacity & (capacity - 1):
            raise AssertionError("capacity must be positive and a power of 2.")
        self._capacity = capacity
        self._t = backend if backend is not None else _NativeTree(capacity)

    @property
    def capacity(self):
        return self._capacity

    # batch leaf write (keys are unique per call: they come from a dict)
    def set_many(self, idx, val):
        # the reference asserts 0 <= idx < capacity per write (segment_tree.py:71); the kernels take the indices as they come
        if len(idx) and not (0 <= min(idx) and max(idx) < self._capacity):
            raise AssertionError("index out of range")
        self._t.set(idx, val)

    def __setitem__(self, idx, val):
        if not (0 <= idx < self._capacity):
            raise AssertionError("index out of range")
        self._t.set([int(idx)], [float(val)])

    def __getitem__(self, idx):
        if not (0 <= idx < self._capacity):
            raise AssertionError("index out of range")
        return float(self._t.values()[self._capacity + idx].item())

    def get_many(self, idx):
        v = self._t.values()
        i = torch.as_tensor(idx, dtype=torch.int64, device=v.device) + self._capacity
        return v[i]

    def sum(self, start=0, end=None):
        """arr[start] + ... + arr[end-1] with the reference's association order"""
        if end is None:
            end = self._capacity
        if end < 0:
            end += self._capacity
        return float(self._t.sum(start, end).item())

    def find_prefixsum_idx(self, prefixsum):
        return int(self._t.find([float(prefixsum)])[0].item())

    def find_many(self, masses):
        return self._t.find(masses)
