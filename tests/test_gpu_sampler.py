"""Philox k-neighbour sampler, to_block and the RBR subset draw on the GPU vs the oracle: index
lists must match bit-for-bit under the shared counter-RNG stream."""
import numpy as np
import pytest
import torch

from oracle.graph import in_csr
from oracle import sampler as osamp

pytestmark = pytest.mark.gpu


def _graph(V, E, seed, isolated=0):
    import ogl_b200
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, V - isolated + 1) ** 1.1
    w /= w.sum()
    src = rng.choice(V - isolated, size=E, p=w).astype(np.int64)
    dst = rng.choice(V - isolated, size=E, p=w).astype(np.int64)
    g = ogl_b200.native.Graph(V, 2 * E)
    g.insert_vertices(V)
    g.insert_edges(torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda(), symmetric=True)
    es, ed = np.concatenate([src, dst]), np.concatenate([dst, src])
    return g, in_csr(es, ed, V)


@pytest.mark.parametrize("fanout", [1, 3, 10, 25, 45])
def test_sample_neighbors_bit_exact(fanout):
    import ogl_b200
    V = 2000
    g, (ip, ix, ei) = _graph(V, 30000, seed=2, isolated=50)     # last 50 vertices have degree 0
    rng = np.random.default_rng(fanout)
    dst = np.concatenate([rng.integers(0, V, 700), np.arange(V - 5, V)]).astype(np.int64)
    for step, hop in ((0, 0), (7, 1), (123456, 0)):
        src_d, eid_d = ogl_b200.native.sample_neighbors(g, dst, fanout, seed=0xC0FFEE1234, step=step, hop=hop)
        rs, re = osamp.sample_neighbors(ip, ix, ei, dst, fanout, 0xC0FFEE1234, step, hop)
        assert np.array_equal(src_d.cpu().numpy().astype(np.int64), rs)
        assert np.array_equal(eid_d.cpu().numpy(), re)
    assert (rs.reshape(-1, fanout)[-5:] == -1).all()


def test_sample_is_uniform_over_the_row():
    """size-independent property: picks of a fixed row over many steps cover its neighbour list uniformly"""
    import ogl_b200
    V = 64
    src = np.arange(1, 41, dtype=np.int64)
    dst = np.zeros(40, dtype=np.int64)
    g = ogl_b200.native.Graph(V, 128)
    g.insert_vertices(V)
    g.insert_edges(torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda(), symmetric=False)
    counts = np.zeros(41)
    for step in range(50):
        s, _ = ogl_b200.native.sample_neighbors(g, np.zeros(256, dtype=np.int64), 10, seed=9, step=step, hop=0)
        counts += np.bincount(s.cpu().numpy(), minlength=41)
    assert counts[0] == 0 and counts[1:].min() > 0.85 * counts[1:].mean() and counts[1:].max() < 1.15 * counts[1:].mean()


@pytest.mark.parametrize("fanouts,n_seeds", [([10, 10], 64), ([25, 10], 300), ([4], 17), ([5, 3, 2], 40)])
def test_plan_minibatch_matches_oracle(fanouts, n_seeds):
    import ogl_b200
    V = 3000
    g, (ip, ix, ei) = _graph(V, 20000, seed=4, isolated=100)
    rng = np.random.default_rng(1)
    seeds = rng.permutation(V)[:n_seeds].astype(np.int64)
    L = len(fanouts)
    plan = ogl_b200.native.Plan([8] * (L + 1), fanouts, 512, V, mode=ogl_b200.OGL_F32, seed=77)
    for step in (0, 5):
        plan.set_step(step)
        plan.sample(g, torch.as_tensor(seeds).cuda())
        input_nodes, blocks = osamp.sample_blocks(ip, ix, ei, seeds, fanouts, 77, step)
        assert np.array_equal(plan.level_nodes(L).cpu().numpy(), input_nodes)
        for hop in range(L):
            b = blocks[L - 1 - hop]                       # oracle lists blocks input layer first
            assert np.array_equal(plan.level_nodes(hop).cpu().numpy(), b["dst_nodes"])
            assert np.array_equal(plan.level_nodes(hop + 1).cpu().numpy(), b["src_nodes"])
            lid, gsrc, eid, f = plan.block_edges(hop)
            assert f == fanouts[hop]
            assert np.array_equal(lid.cpu().numpy(), b["edge_src"])
            assert np.array_equal(gsrc.cpu().numpy(), b["edge_src_global"])
            assert np.array_equal(eid.cpu().numpy(), b["edge_eid"])
        # the fast (vectorised) oracle used by the CPU baseline agrees with the loop oracle
        in2, blocks2 = osamp.sample_blocks(ip, ix, ei, seeds, fanouts, 77, step, fast=True)
        assert np.array_equal(in2, input_nodes)
        assert all(np.array_equal(a["edge_src"], b["edge_src"]) for a, b in zip(blocks, blocks2))


def test_dataloader_surface_yields_dgl_style_blocks():
    import ogl_b200
    from ogl_b200.sampling import MultiLayerNeighborSampler, NodeDataLoader, NID
    V = 500
    rng = np.random.default_rng(8)
    src, dst = rng.integers(0, V, 4000), rng.integers(0, V, 4000)
    dg = ogl_b200.DeviceGraph(V, 8000, 6)
    dg.add_nodes(V, {"feat": np.zeros((V, 6), np.float32), "target": np.zeros((V, 1), np.int64)})
    dg.add_edges(src, dst, symmetric=True)
    ip, ix, ei = in_csr(np.concatenate([src, dst]), np.concatenate([dst, src]), V)
    sampler = MultiLayerNeighborSampler([7, 7], replace=True, return_eids=True)
    nids = np.arange(0, 90)
    loader = NodeDataLoader(dg, nids, sampler, batch_size=40, shuffle=False, drop_last=False, num_workers=0)
    assert len(loader) == 3
    seen = 0
    for step, (input_nodes, seeds, blocks) in enumerate(loader):
        assert len(blocks) == 2 and blocks[0].is_block
        assert seeds.cpu().tolist() == nids[seen:seen + 40].tolist()
        seen += seeds.numel()
        assert torch.equal(blocks[0].srcdata[NID], input_nodes)
        assert torch.equal(blocks[1].dstdata[NID], seeds)
        assert blocks[1].number_of_dst_nodes() == seeds.numel()
        # every block edge is a real in-edge of its destination
        for b in blocks:
            s_l, d_l = b.edges()
            gs = b.srcdata[NID][s_l].cpu().numpy()
            gd = b.dstdata[NID][d_l].cpu().numpy()
            eid = b.edata["_ID"].cpu().numpy()
            es, ed = np.concatenate([src, dst]), np.concatenate([dst, src])
            assert np.array_equal(es[eid], gs) and np.array_equal(ed[eid], gd)


@pytest.mark.parametrize("n_pop,n", [(10, 10), (1000, 32), (150000, 1024), (3, 1), (65536, 65536), (100001, 5000)])
def test_draw_uniform_subset(n_pop, n):
    import ogl_b200
    for counter in (1, 2):
        got = ogl_b200.native.draw_uniform(n_pop, n, seed=42, counter=counter).cpu().numpy()
        assert len(np.unique(got)) == n and got.min() >= 0 and got.max() < n_pop
        if n < n_pop:
            ref = osamp.draw_uniform_subset(n_pop, n, 42, counter)
            assert np.array_equal(got, ref)
        else:
            assert np.array_equal(np.sort(got), np.arange(n_pop))


def test_duplicate_and_out_of_range_seeds_are_safe():
    """DGL accepts duplicate seeds (every copy is a destination row of its own); on a small graph the v_cap clamp of the frontier
    size must leave room for them.  Seed ids outside the graph raise the plan's error flag instead of reading out of bounds."""
    import ogl_b200
    rng = np.random.default_rng(3)
    V, E = 300, 4000
    g = ogl_b200.native.Graph(V, 2 * E)
    g.insert_vertices(V)
    g.insert_edges(torch.as_tensor(rng.integers(0, V, E)).cuda(), torch.as_tensor(rng.integers(0, V, E)).cuda(), symmetric=True)
    plan = ogl_b200.native.Plan([8, 8, 8], [45, 45], 1024, V, mode=ogl_b200.OGL_F32, seed=1)
    seeds = torch.as_tensor(np.concatenate([np.arange(V), rng.integers(0, V, 1024 - V)]).astype(np.int64)).cuda()   # 724 duplicates
    plan.sample(g, seeds)
    torch.cuda.synchronize()
    n0, n1, n2 = (plan.level_nodes(lv).numel() for lv in range(3))
    assert n0 == 1024 and n1 >= 1024 and n2 >= n1
    assert torch.equal(plan.level_nodes(1)[:1024].long(), seeds)
    lid = plan.block_edges(1)[0]
    assert int(lid.max()) < n2 and plan.error_flags() == 0
    bad = seeds.clone()
    bad[5] = V + 7
    bad[9] = -3
    plan.sample(g, bad)
    assert plan.error_flags() == 1 and plan.error_flags() == 0
