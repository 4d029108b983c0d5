"""Philox4x32-10 counter-based RNG (Salmon et al., SC'11) in numpy -- oracle only.

The reference's sampler RNG is DGL's thread-local std::mt19937 (not reproducible;
call site train/graphsage/pytorch/model.py:128).  The B200 build replaces it with
this counter RNG; this file is the bit-exact CPU statement of the device code in
online-gnn-learning_b200/csrc/philox.cuh.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
S32 = np.uint64(32)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32.  Counters are broadcastable uint32 arrays, keys are
    python ints.  Returns a tuple of four uint32 arrays."""
    c0, c1, c2, c3 = np.broadcast_arrays(*(np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3)))
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for r in range(rounds):
        p0 = M0 * c0          # 32x32 -> 64 (fits in uint64)
        p1 = M1 * c2
        hi0, lo0 = p0 >> S32, p0 & MASK
        hi1, lo1 = p1 >> S32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)), lo1, (hi0 ^ c3 ^ np.uint64(k1)), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def split_seed(seed):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, seed >> 32


def mulhi32(w, n):
    """Unbiased-enough range reduction used everywhere on the path:
    floor(w * n / 2**32) for a 32-bit word w and n < 2**32."""
    return ((np.asarray(w, dtype=np.uint64) * np.asarray(n, dtype=np.uint64)) >> S32).astype(np.int64)
