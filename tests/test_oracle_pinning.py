"""Pin the oracle against fixtures produced by the reference's OWN code
(tests/golden/make_golden.py ran /root/reference/train/{prioritized_replay,graph})."""
import numpy as np
from oracle.sumtree import SumTree, PrioritizedBufferOracle
from oracle.graph import EdgeStreamOracle, VertexStreamOracle
from oracle.philox import philox4x32


def test_philox_random123_kat():
    h = lambda t: [int(x) for x in t]
    assert h(philox4x32(0, 0, 0, 0, 0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert h(philox4x32(*[0xffffffff] * 6)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert h(philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_sumtree_vs_reference(golden):
    g = golden("replay_tree")
    t = SumTree(int(g["cap"]))
    # sequential semantics: feed one by one AND as a batch, both must equal the reference
    for i, v in zip(g["idx"].tolist(), g["val"].tolist()):
        t.set([i], [v])
    assert np.array_equal(t.value, g["value"])
    tb = SumTree(int(g["cap"]))
    tb.set(g["idx"], g["val"])
    assert np.array_equal(tb.value, g["value"])
    for (a, b), s in zip(g["ranges"].tolist(), g["sums"].tolist()):
        assert t.sum(a, b) == s
    assert np.array_equal(t.find_prefixsum_idx(g["masses"]), g["found"])
    t8 = SumTree(8)
    t8.set(range(6), [0.5, 1, 0.25, 2, 0, 3])
    assert [t8.sum(), t8.sum(0, 5), float(t8.find_prefixsum_idx(1.6)[0])] == g["t8"].tolist() == [6.75, 3.75, 2.0]


def test_prioritized_buffer_vs_reference(golden):
    g = golden("replay_buffer")
    buf = PrioritizedBufferOracle(int(g["size"]), float(g["alpha"]), float(g["max_p"]), float(g["min_p"]))
    buf.add_all(dict(zip(g["nodes1"].tolist(), g["pri1"].tolist())))
    assert np.array_equal(buf.tree.value, g["leaves1"])
    buf.update_priorities(dict(zip(g["upd_nodes"].tolist(), g["upd_pri"].tolist())))
    assert np.array_equal(buf.tree.value, g["leaves2"])
    assert [buf.min_val, buf.max_val, buf._min_priority, buf._max_priority] == g["minmax"].tolist()
    buf.add_all(dict(zip(g["nodes2"].tolist(), [float(g["p2"])] * len(g["nodes2"]))))
    assert np.array_equal(buf.tree.value, g["leaves3"])
    assert buf.storage == g["storage"].tolist()
    n = int(g["draw_n"])
    u = g["draw_uniforms"].tolist()
    assert buf.tree.sum(0, len(buf) - 1) == float(g["p_total"])
    res = buf.sample_proportional(n, u[:n], u[n:], g["draw_randints"].tolist())
    assert sorted(res) == g["draw_result"].tolist()


def test_edge_stream_vs_reference(golden):
    g = golden("edge_stream")
    o = EdgeStreamOracle(g["src"], g["dst"], int(g["snapshots"]))
    n_nodes, n_edges, newv = [o.n_vertices], [len(o.e_src)], [list(o.new_vertices)]
    for _ in range(10):
        o.evolve()
        n_nodes.append(o.n_vertices); n_edges.append(len(o.e_src)); newv.append(list(o.new_vertices))
    assert n_nodes == g["n_nodes"].tolist() and n_edges == g["n_edges"].tolist()
    assert np.array_equal(o.e_src, g["log_src"]) and np.array_equal(o.e_dst, g["log_dst"])
    assert [len(x) for x in newv] == g["newv_len"].tolist()
    assert [v for x in newv for v in x] == g["newv_flat"].tolist()
    # feature rows follow vertex ids (dense first-appearance relabelling precondition)
    assert np.array_equal(g["feat"][:, 0], np.arange(o.n_vertices) * 3.0)
    assert np.array_equal(o.touched(3), g["added3_vertices"])
    indptr, indices, eids = o.csr()
    assert indptr[-1] == len(o.e_src)
    for v in (0, 1, 5, o.n_vertices - 1):
        e = eids[indptr[v]:indptr[v + 1]]
        assert np.all(np.diff(e) > 0) and np.all(o.e_dst[e] == v)
        assert np.array_equal(o.e_src[e], indices[indptr[v]:indptr[v + 1]])


def test_vertex_stream_vs_reference(golden):
    g = golden("vertex_stream")
    V = int(g["V"])
    ts_v, ts_t = g["ts_vertex"], g["ts_time"]
    order = ts_v[np.argsort(ts_t, kind="stable")]       # list.sort(key=timestamp) is stable
    o = VertexStreamOracle(np.zeros(0, int), np.zeros(0, int), V, order, int(g["snapshots"]))
    assert len(o) == int(g["n_chunks"]) == int(g["len_graph"])
    assert np.array_equal(np.concatenate(o.chunks), g["chunk_flat"])
    assert [len(c) for c in o.chunks] == g["chunk_len"].tolist()
    act = [o.n_active()]
    for _ in range(4):
        o.evolve(); act.append(o.n_active())
    assert act == g["n_active"].tolist()
    assert np.array_equal(o.subgraph_to_original(), g["s2o"])
    assert np.array_equal(o.rank[g["probe"]], g["o2s"])


def inference_fixture(g):
    """(feat, params, requests, per-request expectations) of tests/golden/inference_stream.npz"""
    params = {k[len("param_"):]: g[k] for k in g.files if k.startswith("param_")}
    e0 = np.concatenate([[0], np.cumsum(g["n_edges"])])
    p0 = np.concatenate([[0], np.cumsum(g["nP"])])
    s0 = np.concatenate([[0], np.cumsum(g["nS"])])
    reqs = []
    for r in range(len(g["n_edges"])):
        reqs.append(dict(pairs=g["edges"][e0[r]:e0[r + 1]].tolist(), P=g["P"][p0[r]:p0[r + 1]].tolist(), S=g["S"][s0[r]:s0[r + 1]].tolist(),
                         out=g["out"][p0[r]:p0[r + 1]].tolist(), n_nodes=int(g["n_nodes"][r]),
                         caches={k: g["cache_" + k][r] for k in ("h0proj", "neigh0", "h1", "h1proj", "neigh1", "h2")}))
    return g["feat"], params, reqs


def test_cached_inference_oracle_vs_reference_handler(golden):
    """oracle/inference.py against the reference's own `inference()` method run request by request (make_golden_inference.py)"""
    from oracle.inference import CachedInferenceOracle
    feat, params, reqs = inference_fixture(golden("inference_stream"))
    o = CachedInferenceOracle(feat, params)
    assert len(reqs) == 40
    for r, q in enumerate(reqs):
        P, classes = o.request(q["pairs"])
        assert o.n == q["n_nodes"]
        assert P == q["P"], "request %d: answered vertices (and their order)" % r
        assert o.last_sets[2] == q["S"], "request %d: layer-1 vertex set" % r
        assert classes == q["out"], "request %d: predicted classes" % r
        for k, want in q["caches"].items():
            got = np.zeros_like(want)
            got[:o.n] = o.cache[k]
            np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-6, err_msg="request %d cache %s" % (r, k))
