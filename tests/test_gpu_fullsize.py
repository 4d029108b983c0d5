"""Size-independent properties at a BASELINE-scale shape (Reddit-shaped vertex count, tens of millions of directed
edges, power-law hubs) where the CPU oracle is too slow to be the checker: CSR invariants after streaming inserts,
sampler picks are real in-edges, to_block is a bijection onto the frontier, the sum tree conserves mass."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def big_graph():
    import ogl_b200
    V, E = 232965, 12_000_000                       # 24 M directed edges
    g = torch.Generator(device="cuda").manual_seed(1)
    wt = (torch.arange(V, device="cuda", dtype=torch.float64) + 50.0) ** -0.83
    cdf = torch.cumsum(wt / wt.sum(), 0).float()
    perm = torch.randperm(V, generator=g, device="cuda")
    src = perm[torch.searchsorted(cdf, torch.rand(E, generator=g, device="cuda")).clamp_(max=V - 1)]
    dst = perm[torch.searchsorted(cdf, torch.rand(E, generator=g, device="cuda")).clamp_(max=V - 1)]
    graph = ogl_b200.native.Graph(V, 2 * E)
    graph.insert_vertices(V)
    chunk = 1 << 21
    for a in range(0, E, chunk):                    # streamed in 6 batches, hubs receive thousands of edges per batch
        graph.insert_edges(src[a:a + chunk], dst[a:a + chunk], symmetric=True)
    return graph, src, dst, V, E, chunk


def test_streaming_csr_invariants_at_scale(big_graph):
    graph, src, dst, V, E, chunk = big_graph
    assert graph.num_edges == 2 * E
    indptr, indices, eids = graph.export_csr()
    deg = graph.degrees()
    assert int(indptr[-1]) == 2 * E and torch.equal(indptr[1:] - indptr[:-1], deg)
    # degrees == histogram of the symmetrised destinations
    ref_deg = torch.bincount(torch.cat([dst, src]), minlength=V)
    assert torch.equal(deg, ref_deg)
    # every edge id appears exactly once (checksum + range) ...
    assert int(eids.min()) == 0 and int(eids.max()) == 2 * E - 1
    assert int(eids.sum()) == (2 * E) * (2 * E - 1) // 2
    # ... and is strictly ascending inside its row (canonical order)
    row = torch.repeat_interleave(torch.arange(V, device="cuda"), deg)
    same_row = row[1:] == row[:-1]
    assert bool((eids[1:] > eids[:-1])[same_row].all())
    # the edge with id e is the e-th inserted directed edge: batch b holds forward edges then reverse edges
    b = eids // (2 * chunk)
    off = eids - b * 2 * chunk
    n_b = torch.clamp(torch.tensor(E, device="cuda") - b * chunk, max=chunk)
    fwd = off < n_b
    k = b * chunk + torch.where(fwd, off, off - n_b)
    exp_src = torch.where(fwd, src[k], dst[k])
    exp_dst = torch.where(fwd, dst[k], src[k])
    assert torch.equal(indices, exp_src) and torch.equal(row, exp_dst)
    st = graph.stats()
    assert st["relocations"] > 0


def test_sampler_and_to_block_properties_at_scale(big_graph):
    import ogl_b200
    graph, src, dst, V, E, chunk = big_graph
    indptr, indices, eids = graph.export_csr()
    rows = torch.randint(0, V, (200_000,), generator=torch.Generator(device="cuda").manual_seed(3), device="cuda")
    s, e = ogl_b200.native.sample_neighbors(graph, rows, 25, seed=5, step=1, hop=0)
    s2, e2 = ogl_b200.native.sample_neighbors(graph, rows, 25, seed=5, step=1, hop=0)
    assert torch.equal(s, s2) and torch.equal(e, e2)                       # counter RNG: reproducible
    deg = graph.degrees()[rows]
    s, e = s.view(-1, 25), e.view(-1, 25)
    assert bool(((s >= 0) == (deg > 0)[:, None]).all())
    # every pick is a real in-edge of its row: the picked edge id lies in the row and carries the picked source
    eid_to_pos = torch.empty(2 * E, dtype=torch.int64, device="cuda")
    eid_to_pos[eids] = torch.arange(2 * E, device="cuda")
    ok = e >= 0
    pos = eid_to_pos[e[ok]]
    r = rows[:, None].expand(-1, 25)[ok]
    assert bool(((pos >= indptr[r]) & (pos < indptr[r + 1])).all())
    assert torch.equal(indices[pos], s[ok].long())
    # 2-hop minibatch: level lists are duplicate-free, dst-first, and edge_lid maps every pick onto its source
    plan = ogl_b200.native.Plan([8, 8, 8], [25, 10], 1024, V, mode=ogl_b200.OGL_F32, seed=2)
    seeds = torch.randperm(V, device="cuda")[:1024]
    plan.sample(graph, seeds)
    for hop in range(2):
        d, sn = plan.level_nodes(hop).long(), plan.level_nodes(hop + 1).long()
        lid, gsrc, eid, f = plan.block_edges(hop)
        assert sn.unique().numel() == sn.numel() and torch.equal(sn[:d.numel()], d)
        valid = lid >= 0
        assert torch.equal(sn[lid[valid].long()], gsrc[valid].long())
        assert set(gsrc[valid].unique().tolist()) <= set(sn.tolist())


def test_sum_tree_mass_conservation_at_reference_capacity():
    """capacity 2^24 (the reference's 10 M-slot buffer): root == sum of leaves, prefix search inverts the cumulative sum"""
    import ogl_b200
    cap = 1 << 24
    t = ogl_b200.native.SumTree(cap)
    n = 3_000_000
    gen = torch.Generator(device="cuda").manual_seed(0)
    idx = torch.randperm(cap, generator=gen, device="cuda")[:n]
    val = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64) + 0.01
    t.set(idx, val)
    vals = t.values()
    leaves = vals[cap:]
    assert abs(float(vals[1]) - float(leaves.sum())) <= 1e-6 * float(vals[1])
    lvl = vals[cap // 2:cap]
    assert torch.equal(lvl, leaves[0::2] + leaves[1::2])                    # parent = left + right, exactly
    mass = torch.rand(4096, generator=gen, device="cuda", dtype=torch.float64) * float(vals[1]) * 0.999
    found = t.find(mass)
    cs = torch.cumsum(leaves, 0)
    ref = torch.searchsorted(cs, mass, right=True)
    assert float((found == ref).double().mean()) > 0.999                   # (cumsum rounding may differ at exact boundaries)
    assert bool((leaves[found] > 0).all())
