"""GPU fp64 sum-tree + prioritized buffer vs fixtures recorded from the reference's own
train/prioritized_replay code (bit-exact) and vs the oracle restatement on larger random cases."""
import random

import numpy as np
import pytest
import torch

from oracle.sumtree import SumTree as OracleTree, PrioritizedBufferOracle

pytestmark = pytest.mark.gpu


def test_sumtree_golden_bit_exact(golden):
    import ogl_b200
    g = golden("replay_tree")
    cap = int(g["cap"])
    t = ogl_b200.native.SumTree(cap)
    # sequential single writes (reference __setitem__ order) ...
    for i, v in zip(g["idx"].tolist(), g["val"].tolist()):
        t.set([i], [v])
    assert np.array_equal(t.values().cpu().numpy()[1:], g["value"][1:])
    for (a, b), s in zip(g["ranges"].tolist(), g["sums"].tolist()):
        assert float(t.sum(a, b).item()) == s
    assert np.array_equal(t.find(g["masses"]).cpu().numpy(), g["found"])
    t8 = ogl_b200.native.SumTree(8)
    t8.set(list(range(6)), [0.5, 1, 0.25, 2, 0, 3])
    assert [float(t8.sum().item()), float(t8.sum(0, 5).item()), int(t8.find([1.6])[0].item())] == [6.75, 3.75, 2]


@pytest.mark.parametrize("cap,n", [(1 << 10, 700), (1 << 16, 40000), (1 << 20, 300000)])
def test_sumtree_batch_update_vs_oracle(cap, n):
    """unique-key batches (what dict-driven updates produce), small-CTA and multi-launch paths"""
    import ogl_b200
    rng = np.random.default_rng(cap)
    t = ogl_b200.native.SumTree(cap)
    o = OracleTree(cap)
    for rnd in range(3):
        idx = rng.permutation(cap)[:n].astype(np.int64)
        val = rng.random(n) ** 4 * 3.0
        t.set(idx, val)
        o.set(idx, val)
        assert np.array_equal(t.values().cpu().numpy()[1:], o.value[1:])
    m = rng.random(5000) * o.value[1]
    assert np.array_equal(t.find(m).cpu().numpy(), o.find_prefixsum_idx(m))
    for lo, hi in ((0, cap), (0, n - 1), (3, 4), (cap // 3, cap // 2 + 5), (5, 5)):
        ref = o.sum(lo, hi) if hi > lo else 0.0
        assert float(t.sum(lo, hi).item()) == ref
    # stratified draw under shared uniforms
    u = rng.random(1024)
    n_items = n
    p_total = o.sum(0, n_items - 1)
    every = p_total / 1024
    mass = np.array([u[i] * every + i * every for i in range(1024)])
    assert np.array_equal(t.sample_stratified(u, n_items).cpu().numpy(), o.find_prefixsum_idx(mass))


def test_prioritized_buffer_golden(golden):
    """the drop-in PrioritizedReplayBuffer over the GPU tree reproduces the reference's leaves, running min/max and draw"""
    from ogl_b200.prioritized_replay.replay_buffer import PrioritizedReplayBuffer
    g = golden("replay_buffer")
    buf = PrioritizedReplayBuffer(int(g["size"]), float(g["alpha"]), float(g["max_p"]), float(g["min_p"]), verbose=False)
    vals = lambda: buf._it_sum._t.values().cpu().numpy()
    buf.add_all(dict(zip(g["nodes1"].tolist(), g["pri1"].tolist())))
    assert np.array_equal(vals()[1:], g["leaves1"][1:])
    buf.update_priorities(dict(zip(g["upd_nodes"].tolist(), g["upd_pri"].tolist())))
    assert np.array_equal(vals()[1:], g["leaves2"][1:])
    assert [buf.get_min_priority(), buf.get_max_priority(), buf._min_priority, buf._max_priority] == g["minmax"].tolist()
    buf.add_all(dict(zip(g["nodes2"].tolist(), [float(g["p2"])] * len(g["nodes2"]))))
    assert np.array_equal(vals()[1:], g["leaves3"][1:])
    assert buf._it_sum.sum(0, len(buf) - 1) == float(g["p_total"])
    random.seed(123)
    res = buf._sample_proportional(int(g["draw_n"]))
    assert sorted(res) == g["draw_result"].tolist()
    got = buf.sample(int(g["draw_n"]))
    assert len(got) == int(g["draw_n"]) and set(got) <= set(g["storage"].tolist())


def test_device_loss_transform_close_to_host_transform():
    """set_from_loss (clip/log/normalise/pow on the GPU) vs the reference's Python-float transform: same running
    min/max (exact) and leaves within 4 ulp (libm log/pow rounding)"""
    import ogl_b200
    rng = np.random.default_rng(0)
    n, cap = 5000, 8192
    loss = (rng.random(n) ** 3 * 12).astype(np.float32)
    loss[:3] = [0.0, 50.0, 1e-9]                          # clipped both ways
    o = PrioritizedBufferOracle(cap, alpha=4.0, max_priority=10.0, min_priority=1e-7)
    o.add_all({i: float(loss[i]) for i in range(n)})
    t = ogl_b200.native.SumTree(cap)
    state = torch.tensor([99999999.0, -1.0, 99999999.0, -1.0], dtype=torch.float64, device="cuda")
    t.set_from_loss(np.arange(n), loss, 1e-7, 10.0, 1e-5, 4.0, state)
    s = state.cpu().tolist()
    assert s[0] == o.min_val and s[1] == o.max_val
    assert abs(s[2] - o._min_priority) <= 1e-15 * abs(o._min_priority) and abs(s[3] - o._max_priority) <= 1e-15 * abs(o._max_priority)
    got = t.values().cpu().numpy()[cap:cap + n]
    ref = o.tree.value[cap:cap + n]
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-300)
    # root consistent with its own leaves (tree invariant holds exactly on the device values)
    chk = OracleTree(cap)
    chk.set(np.arange(n), got)
    assert np.array_equal(chk.value[1:], t.values().cpu().numpy()[1:])


def test_train_test_graph_on_device_matches_reference(golden):
    """TrainTestGraph + DynamicGraphEdge + GPU sum tree end to end against the fixture recorded from the reference"""
    import ogl_b200
    from ogl_b200 import config
    from ogl_b200.graph import train_test_graph as ttg
    g = golden("train_test")
    config.set_faithful(True)
    old = ttg.SIZE_BUFFER
    ttg.SIZE_BUFFER = int(g["size_buffer"])
    try:
        V = int(g["V"])
        labelled = set(g["labelled"].tolist())
        np.random.seed(1)
        random.seed(1)
        feats = np.zeros((V, 4), np.float32)
        targets = np.zeros((V, 1), np.int64)
        dyn = ogl_b200.DynamicGraphEdge(int(g["snapshots"]), labelled)
        dyn.build(feats, targets, edge_timestamps={"src": g["src"], "dst": g["dst"]})
        tt = ttg.TrainTestGraph(dyn, split=0.15, start_prior_alpha=4, end_prior_alpha=50, scale=1, max_priority=10)
        unpad = lambda row: [int(x) for x in row if x >= 0]
        for k in range(8):
            assert len(tt.get_train_set()) == g["train_len"][k] and len(tt.get_test_set()) == g["test_len"][k]
            assert list(tt.draw_random_train_nodes(16)) == unpad(g["rbr"][k])
            pbr = list(tt.draw_priority_train_nodes(16))
            assert [int(x) for x in pbr] == unpad(g["pbr"][k])
            assert list(tt.get_new_train_nodes(5)) == unpad(g["newn"][k])
            tt.update_priorities({int(v): 0.1 + 0.01 * (int(v) % 37) for v in pbr})
            assert tt.prior_alpha == g["alpha"][k]
            tt.evolve()
        assert tt.get_train_set() == g["final_train"].tolist() and tt.get_test_set() == g["final_test"].tolist()
        assert tt.dump_priorities(tt.get_train_set()) == g["final_priorities"].tolist()
    finally:
        ttg.SIZE_BUFFER = old
