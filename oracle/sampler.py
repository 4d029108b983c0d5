"""Uniform k-neighbour sampler + to_block -- oracle only (numpy).

Restates dgl.sampling.MultiLayerNeighborSampler([s]*2, replace=True, return_eids=True)
+ NodeDataLoader as called at train/graphsage/pytorch/model.py:44-47,128-131
[DGL semantics recalled: hops are sampled from the seeds outwards; with replace=True a
row with in-degree>0 yields exactly `fanout` picks, a row with in-degree 0 yields none;
to_block puts the dst nodes first and then new sources in first-appearance order].
PARITY UNPINNED against DGL (absent); the RNG stream is this repo's specification:

    word  = philox4x32_10(counter=(j>>2, row_position, hop, step), key=seed)[j & 3]
    pick  = (word * deg) >> 32           # index into the row's in-neighbour list
"""
import numpy as np
from .philox import philox4x32, split_seed, mulhi32


def sample_neighbors(indptr, indices, eids, dst_nodes, fanout, seed, step, hop):
    """Returns (picked_src[int64 n*fanout], picked_eid[int64 n*fanout]) in (row, j) order;
    rows with degree 0 hold -1 in all fanout slots."""
    dst_nodes = np.asarray(dst_nodes, dtype=np.int64)
    n = len(dst_nodes)
    start = indptr[dst_nodes]
    deg = indptr[dst_nodes + 1] - start
    j = np.arange(fanout, dtype=np.int64)[None, :]
    row = np.arange(n, dtype=np.int64)[:, None]
    k0, k1 = split_seed(seed)
    words = philox4x32(j >> 2, row, hop, step, k0, k1)
    lane = np.broadcast_to(j & 3, (n, fanout))
    w = np.choose(lane, [np.broadcast_to(x, (n, fanout)) for x in words])
    r = mulhi32(w, np.broadcast_to(deg[:, None], (n, fanout)))
    pos = start[:, None] + r
    has = np.broadcast_to((deg > 0)[:, None], (n, fanout))
    pos = np.where(has, pos, 0)
    if len(indices) == 0:
        src = np.full((n, fanout), -1, dtype=np.int64)
        eid = np.full((n, fanout), -1, dtype=np.int64)
    else:
        src = np.where(has, indices[pos], -1)
        eid = np.where(has, eids[pos], -1)
    return src.reshape(-1), eid.reshape(-1)


def to_block(dst_nodes, picked_src):
    """Returns (src_nodes[int64 n_src], edge_src_local[int64 n_dst*fanout]) where src_nodes
    = dst_nodes followed by first-appearance-ordered new sources, and edge_src_local maps
    every slot to its position in src_nodes (-1 for empty slots)."""
    dst_nodes = np.asarray(dst_nodes, dtype=np.int64)
    lid = {}
    for i, g in enumerate(dst_nodes.tolist()):
        lid.setdefault(g, i)
    nodes = dst_nodes.tolist()
    out = np.full(len(picked_src), -1, dtype=np.int64)
    for p, g in enumerate(np.asarray(picked_src).tolist()):
        if g < 0:
            continue
        if g not in lid:
            lid[g] = len(nodes)
            nodes.append(g)
        out[p] = lid[g]
    return np.asarray(nodes, dtype=np.int64), out


def to_block_fast(dst_nodes, picked_src):
    """Vectorised to_block (same result; used by the CPU baseline at full sizes)."""
    dst_nodes = np.asarray(dst_nodes, dtype=np.int64)
    picked_src = np.asarray(picked_src, dtype=np.int64)
    n_dst = len(dst_nodes)
    valid = picked_src >= 0
    allv = np.concatenate([dst_nodes, picked_src[valid]])
    uniq, first = np.unique(allv, return_index=True)
    order = np.argsort(first, kind="stable")          # first-appearance order
    nodes = uniq[order]
    rank_of_uniq = np.empty(len(uniq), dtype=np.int64)
    rank_of_uniq[order] = np.arange(len(uniq))
    out = np.full(len(picked_src), -1, dtype=np.int64)
    out[valid] = rank_of_uniq[np.searchsorted(uniq, picked_src[valid])]
    assert np.array_equal(nodes[:n_dst], dst_nodes), "dst nodes must be unique"
    return nodes, out


def sample_blocks(indptr, indices, eids, seeds, fanouts, seed, step, fast=False):
    """Two-hop (len(fanouts)-hop) minibatch.  ``fanouts`` is listed seeds-hop first, i.e.
    fanouts[0] applies to the seeds (hop 0), fanouts[1] to the outer hop (hop 1) -- DGL
    lists them input-layer first, the reference always passes equal values
    (pytorch/model.py:128).  Returns (input_nodes, blocks) with blocks ordered input layer
    first like DGL; each block is a dict(dst_nodes, src_nodes, edge_src, edge_eid, fanout)."""
    tb = to_block_fast if fast else to_block
    blocks = []
    dst = np.asarray(seeds, dtype=np.int64)
    for hop, s in enumerate(fanouts):
        psrc, peid = sample_neighbors(indptr, indices, eids, dst, s, seed, step, hop)
        nodes, esrc = tb(dst, psrc)
        blocks.insert(0, dict(dst_nodes=dst, src_nodes=nodes, edge_src=esrc, edge_eid=peid,
                              edge_src_global=psrc, fanout=s))
        dst = nodes
    return dst, blocks


# ---- uniform n-subset draw (RBR), device twin: csrc/replay.cu:feistel_draw -------------
def _feistel(x, bits, seed, counter):
    """Keyed 4-round balanced Feistel permutation of [0, 2**bits) (bits even)."""
    half = bits // 2
    mask = np.uint64((1 << half) - 1)
    k0, k1 = split_seed(seed)
    x = np.asarray(x, dtype=np.uint64)
    l, r = x >> np.uint64(half), x & mask
    for rnd in range(4):
        f = philox4x32(r, rnd, counter, 0x5EED, k0, k1)[0].astype(np.uint64) & mask
        l, r = r, l ^ f
    return (l << np.uint64(half)) | r


def draw_uniform_subset(n_pop, n, seed, counter):
    """Indices (into the population list) of a uniform n-subset without replacement, in
    draw order: the first n values pi(i) < n_pop for i = 0, 1, 2, ... where pi is a keyed
    Feistel bijection on the next even power of two >= n_pop.  Replaces the O(n_pop)
    in-place random.shuffle of train/graph/train_test_graph.py:210-216."""
    if n >= n_pop:
        return np.arange(n_pop, dtype=np.int64)
    bits = max(2, int(n_pop - 1).bit_length())
    bits += bits & 1
    out = []
    i = 0
    while len(out) < n:
        cand = _feistel(np.arange(i, i + 4 * n, dtype=np.uint64), bits, seed, counter).astype(np.int64)
        out.extend(cand[cand < n_pop].tolist())
        i += 4 * n
    return np.asarray(out[:n], dtype=np.int64)
