"""Streaming CSR on the GPU vs the oracle's canonical graph (bit-exact: indptr, indices, edge ids,
degrees), through the C ABI.  Edge cases: duplicates, self loops, empty batches, hub rows with long
per-batch tails, relocation / pool rebuild, multi-chunk symmetric inserts, bad ids, capacity."""
import numpy as np
import pytest
import torch

from oracle.graph import EdgeStreamOracle, VertexStreamOracle, in_csr

pytestmark = pytest.mark.gpu


def _csr_np(g):
    ip, ix, ei = g.export_csr()
    torch.cuda.synchronize()
    return ip.cpu().numpy(), ix.cpu().numpy(), ei.cpu().numpy()


def _check(g, e_src, e_dst, V):
    ip, ix, ei = _csr_np(g)
    rp, rx, re = in_csr(e_src, e_dst, V)
    assert np.array_equal(ip, rp)
    assert np.array_equal(ix, rx)
    assert np.array_equal(ei, re)
    assert np.array_equal(g.degrees().cpu().numpy(), np.diff(rp))


def _powerlaw_stream(n_edges, V, seed, gamma=1.2):
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, V + 1) ** gamma
    w /= w.sum()
    src = rng.choice(V, size=n_edges, p=w)
    dst = rng.choice(V, size=n_edges, p=w)
    return src.astype(np.int64), dst.astype(np.int64)


def test_edge_stream_golden(golden):
    import ogl_b200
    g = golden("edge_stream")
    o = EdgeStreamOracle(g["src"], g["dst"], int(g["snapshots"]))
    dg = ogl_b200.native.Graph(int(g["src"].max() + g["dst"].max()) + 2, 2 * len(g["src"]))
    eps = o.edges_per_snapshot
    seen = 0
    for k in range(11):
        if k:
            o.evolve()
        dg.insert_vertices(o.n_vertices - seen)
        seen = o.n_vertices
        s, d = g["src"][k * eps:(k + 1) * eps], g["dst"][k * eps:(k + 1) * eps]
        dg.insert_edges(torch.as_tensor(s).cuda(), torch.as_tensor(d).cuda(), symmetric=True)
        assert dg.num_vertices == g["n_nodes"][k] and dg.num_edges == g["n_edges"][k]
        _check(dg, o.e_src, o.e_dst, o.n_vertices)
    assert np.array_equal(o.e_src, g["log_src"])


@pytest.mark.parametrize("host", [False, True])
def test_powerlaw_stream_with_hubs_and_relocation(host):
    import ogl_b200
    V, E, B = 3000, 120000, 6000
    src, dst = _powerlaw_stream(E, V, seed=5)
    src[::97] = dst[::97]                      # self loops
    src[1::211], dst[1::211] = src[0], dst[0]  # exact duplicates
    dg = ogl_b200.native.Graph(V, 64)          # tiny pool: forces rebuilds + relocations
    dg.insert_vertices(V)
    es, ed = np.zeros(0, np.int64), np.zeros(0, np.int64)
    for a in range(0, E, B):
        s, d = src[a:a + B], dst[a:a + B]
        if host:
            dg.insert_edges(s, d, symmetric=True)
        else:
            dg.insert_edges(torch.as_tensor(s).cuda(), torch.as_tensor(d).cuda(), symmetric=True)
        es = np.concatenate([es, s, d])
        ed = np.concatenate([ed, d, s])
        if a in (0, 5 * B, E - B):
            _check(dg, es, ed, V)
    st = dg.stats()
    assert st["relocations"] > 0 and st["compactions"] > 0
    dg.compact()
    _check(dg, es, ed, V)
    assert dg.stats()["pool_used"] <= dg.stats()["pool_cap"]


def test_giant_tail_single_batch():
    """one batch where one row receives > 2048 edges (CTA-wide tail ordering) and several > 32"""
    import ogl_b200
    V = 500
    rng = np.random.default_rng(3)
    n = 40000
    dst = np.where(rng.random(n) < 0.3, 7, rng.integers(0, 40, n)).astype(np.int64)
    src = rng.integers(0, V, n).astype(np.int64)
    dg = ogl_b200.native.Graph(V, 4 * n)
    dg.insert_vertices(V)
    dg.insert_edges(torch.as_tensor(src).cuda(), torch.as_tensor(dst).cuda(), symmetric=False)
    _check(dg, src, dst, V)
    dg.insert_edges(torch.as_tensor(src[:100]).cuda(), torch.as_tensor(dst[:100]).cuda(), symmetric=True)
    _check(dg, np.concatenate([src, src[:100], dst[:100]]), np.concatenate([dst, dst[:100], src[:100]]), V)


def test_empty_and_ragged_inputs():
    import ogl_b200
    dg = ogl_b200.native.Graph(10, 16)
    ip, ix, ei = _csr_np(dg)
    assert ip.tolist() == [0] and len(ix) == 0
    dg.insert_vertices(4)
    dg.insert_edges(torch.zeros(0, dtype=torch.int64).cuda(), torch.zeros(0, dtype=torch.int64).cuda())
    ip, ix, ei = _csr_np(dg)
    assert ip.tolist() == [0, 0, 0, 0, 0]
    dg.insert_edges(torch.tensor([0, 0, 3]).cuda(), torch.tensor([1, 1, 3]).cuda(), symmetric=True)
    _check(dg, np.array([0, 0, 3, 1, 1, 3]), np.array([1, 1, 3, 0, 0, 3]), 4)


def test_bad_ids_and_capacity_fail_loudly():
    import ogl_b200
    dg = ogl_b200.native.Graph(8, 16)
    dg.insert_vertices(4)
    with pytest.raises(ogl_b200.OglError):
        dg.insert_edges(torch.tensor([0, 5]).cuda(), torch.tensor([1, 2]).cuda())
    with pytest.raises(ogl_b200.OglError):
        dg.insert_edges(torch.tensor([-1]).cuda(), torch.tensor([1]).cuda())
    # the graph stays usable after a rejected batch
    dg.insert_edges(torch.tensor([0, 2]).cuda(), torch.tensor([1, 3]).cuda(), symmetric=True)
    _check(dg, np.array([0, 2, 1, 3]), np.array([1, 3, 0, 2]), 4)
    with pytest.raises(ogl_b200.OglError):
        dg.insert_vertices(100)


def test_vertex_stream_prefix_subgraph():
    import ogl_b200
    V, E = 800, 6000
    rng = np.random.default_rng(11)
    u, v = rng.integers(0, V, E), rng.integers(0, V, E)
    p_src, p_dst = np.concatenate([u, v]), np.concatenate([v, u])
    order = rng.permutation(V).astype(np.int64)
    o = VertexStreamOracle(p_src, p_dst, V, order, snapshots=20)
    rank = o.rank
    tmp = ogl_b200.native.Graph(V, 2 * E)
    tmp.insert_vertices(V)
    tmp.insert_edges(torch.as_tensor(rank[p_src]).cuda(), torch.as_tensor(rank[p_dst]).cuda(), symmetric=False)
    ip, ix, ei = tmp.export_csr()
    dg = ogl_b200.native.Graph(V, 2 * E)
    dg.load_parent(ip, ix, ei)
    for step in range(6):
        n = o.n_active()
        dg.set_active_prefix(n)
        rp, rx, re = o.csr()
        gp, gx, ge = _csr_np(dg)
        assert dg.num_vertices == n and dg.num_edges == rp[-1]
        assert np.array_equal(gp, rp) and np.array_equal(gx, rx) and np.array_equal(ge, re)
        o.evolve()
    dg.set_active_prefix(V)
    gp, gx, ge = _csr_np(dg)
    rp, rx, re = in_csr(rank[p_src], rank[p_dst], V)
    assert np.array_equal(gp, rp) and np.array_equal(gx, rx) and np.array_equal(ge, re)
    dg.set_active_prefix(0)
    assert dg.num_edges == 0


def test_dynamic_graph_edge_mirror_matches_reference_fixture(golden):
    """the drop-in DynamicGraphEdge (Python mirror over the kernels) against the fixture recorded from the
    reference's own dynamic_graph_edge.py"""
    import ogl_b200
    g = golden("edge_stream")
    V = int(g["n_nodes"][-1])
    Vall = int(max(g["src"].max(), g["dst"].max())) + 1
    feats = np.zeros((Vall, 4), dtype=np.float32)
    feats[:, 0] = np.arange(Vall) * 3.0
    targets = (np.arange(Vall) % 3).reshape(-1, 1)
    dyn = ogl_b200.DynamicGraphEdge(int(g["snapshots"]), set(range(Vall)))
    dyn.build(feats, targets, edge_timestamps={"src": g["src"], "dst": g["dst"]})
    n_nodes, n_edges, newv = [len(dyn.get_graph())], [dyn.get_graph().number_of_edges()], [sorted(dyn.new_vertices)]
    for _ in range(10):
        dyn.evolve()
        n_nodes.append(len(dyn.get_graph()))
        n_edges.append(dyn.get_graph().number_of_edges())
        newv.append(sorted(dyn.new_vertices))
    assert n_nodes == g["n_nodes"].tolist() and n_edges == g["n_edges"].tolist()
    assert [len(x) for x in newv] == g["newv_len"].tolist()
    assert [v for x in newv for v in x] == g["newv_flat"].tolist()
    ip, ix, ei = dyn.get_graph().csr()
    rp, rx, re = in_csr(g["log_src"], g["log_dst"], V)
    assert np.array_equal(ip.cpu().numpy(), rp) and np.array_equal(ix.cpu().numpy(), rx) and np.array_equal(ei.cpu().numpy(), re)
    got = dyn.get_graph().ndata["feat"][:, 0].cpu().numpy()
    assert np.array_equal(got, g["feat"][:, 0])
    added, lab = dyn.get_added_vertices(3)
    assert np.array_equal(np.asarray(sorted(added)), g["added3_vertices"])


def test_predecessors_and_change_propagation():
    """DGL-style predecessors / out_degree on the device graph and the reference's (disabled) 2-hop priority bump"""
    import ogl_b200
    from ogl_b200.graph import train_test_graph as ttg
    rng = np.random.default_rng(4)
    V, E = 300, 900
    src, dst = rng.integers(0, V, E), rng.integers(0, V, E)
    m = np.full(V, -1)
    order = {}
    for a, b in zip(src.tolist(), dst.tolist()):
        for w in (a, b):
            order.setdefault(w, len(order))
    perm = np.array(sorted(order, key=order.get))
    m[perm] = np.arange(len(perm))
    src, dst = m[src], m[dst]
    Vn = len(perm)
    dyn = ogl_b200.DynamicGraphEdge(1, set(range(Vn)))
    dyn.build(np.zeros((Vn, 4), np.float32), np.zeros((Vn, 1), np.int64), edge_timestamps={"src": src, "dst": dst})
    g = dyn.get_graph()
    ip, ix, _ = in_csr(np.concatenate([src, dst]), np.concatenate([dst, src]), Vn)
    for v in (0, 1, 17, Vn - 1):
        assert g.predecessors(v).tolist() == ix[ip[v]:ip[v + 1]].tolist()
        assert g.out_degree(v) == g.in_degree(v) == ip[v + 1] - ip[v]
    old = ttg.SIZE_BUFFER
    ttg.SIZE_BUFFER = 1 << 10
    try:
        tt = ttg.TrainTestGraph(dyn, split=0.15, start_prior_alpha=4, end_prior_alpha=50, scale=1, max_priority=10)
        got = tt._get_affected_nodes(5, depth=2)
    finally:
        ttg.SIZE_BUFFER = old
    # restatement of the reference loop over the oracle CSR
    nbrs = {5: 1}
    for _ in range(2):
        tmp = {}
        for k, val in nbrs.items():
            for nb in ix[ip[k]:ip[k + 1]].tolist():
                w = (1 / (ip[nb + 1] - ip[nb])) * val
                tmp[nb] = min(tmp[nb] + w, 1) if nb in tmp else w
        for k, val in tmp.items():
            nbrs[k] = max(nbrs[k], val) if k in nbrs else val
    assert got == nbrs


def test_fused_snapshot_insert_equals_general_path(monkeypatch):
    """snapshot-sized batches run through ONE cooperative kernel (k_insert_fused); the result must be the general multi-kernel
    path's, bit for bit: same CSR, same edge ids -- including a hub row of > 2^14 edges that relocates inside a small batch (queued
    copy swept by all CTAs), tails of every size class, a bad id (nothing changes) and pool exhaustion (fallback + rebuild)"""
    import ogl_b200
    V = 4000
    rng = np.random.default_rng(11)
    hub = 5
    # a big first batch makes vertex 5 a hub (general path), then 60 snapshots of 3000 stream edges, 20 % of them on the hub
    first = (rng.integers(0, V, 70000).astype(np.int64), np.full(70000, hub, dtype=np.int64))
    snaps = []
    for k in range(60):
        s = rng.integers(0, V, 3000).astype(np.int64)
        d = np.where(rng.random(3000) < 0.2, hub, (rng.random(3000) ** 3 * V).astype(np.int64)).astype(np.int64)
        snaps.append((s, d))
    out = []
    for fused in ("1", "0"):
        monkeypatch.setenv("OGL_INSERT_FUSED", fused)
        dg = ogl_b200.native.Graph(V, 600000)
        dg.insert_vertices(V)
        dg.insert_edges(torch.as_tensor(first[0]).cuda(), torch.as_tensor(first[1]).cuda(), symmetric=False)
        launches = ogl_b200.kernel_launches()
        for s, d in snaps:
            dg.insert_edges(torch.as_tensor(s).cuda(), torch.as_tensor(d).cuda(), symmetric=True)
        per_snapshot = (ogl_b200.kernel_launches() - launches) / len(snaps)
        bad = torch.as_tensor(np.array([1, 2, V + 3], dtype=np.int64)).cuda()
        with pytest.raises(Exception):
            dg.insert_edges(bad, bad, symmetric=True)
        s, d = snaps[0]
        dg.insert_edges(torch.as_tensor(s).cuda(), torch.as_tensor(d).cuda(), symmetric=True)      # still usable after the error
        out.append([t.cpu() for t in dg.export_csr()] + [per_snapshot, dg.stats()["relocations"]])
    for a, b in zip(out[0][:3], out[1][:3]):
        assert torch.equal(a, b)
    assert out[0][3] <= 2.1 and out[1][3] >= 9, (out[0][3], out[1][3])          # launches per snapshot: fused vs general
    assert out[0][4] == out[1][4] and out[0][4] > 0
    es = np.concatenate([first[0]] + [np.concatenate([s, d]) for s, d in snaps] + [np.concatenate(snaps[0])])
    ed = np.concatenate([first[1]] + [np.concatenate([d, s]) for s, d in snaps] + [np.concatenate(snaps[0][::-1])])
    dg2 = dg
    _check(dg2, es, ed, V)
