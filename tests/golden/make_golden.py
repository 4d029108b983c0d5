"""Generate the committed golden fixtures by RUNNING THE REFERENCE'S OWN CODE
(/root/reference/train/{prioritized_replay,graph}) in the build container.

    python tests/golden/make_golden.py

Outputs (all small .npz, committed):
  replay_tree.npz     SumSegmentTree set / sum / find_prefixsum_idx vectors
  replay_buffer.npz   PrioritizedReplayBuffer add_all / update_priorities leaves + a
                      _sample_proportional draw with the consumed random stream recorded
  edge_stream.npz     DynamicGraphEdge.build/evolve: directed edge log, vertex counts,
                      new-vertex lists, get_added_vertices(delta)
  vertex_stream.npz   DynamicGraphVertex.build/evolve: active lists + both id maps
  train_test.npz      TrainTestGraph over the edge stream under random.seed(1) /
                      np.random.seed(1): train/test sets and RBR/PBR/new-node draws
  priority_strategies.npz  TrendPriority / HybridPriority outputs over 12 batches
"""
import os
import io
import random
import contextlib
import numpy as np
import pandas as pd
import torch

from _ref_import import load_reference

OUT = os.path.dirname(os.path.abspath(__file__))
R = load_reference()


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def gen_tree():
    T = R["segment_tree"].SumSegmentTree
    rng = np.random.default_rng(11)
    cap = 64
    t = T(cap)
    idx = rng.integers(0, 50, 200)
    val = rng.random(200) * 3
    for i, v in zip(idx.tolist(), val.tolist()):
        t[i] = v
    leaves = np.array(t._value, dtype=np.float64)
    ranges = np.array([[0, 64], [0, 49], [3, 17], [5, 6], [10, 64], [31, 33], [0, 1], [7, 50]])
    sums = np.array([t.sum(int(a), int(b)) for a, b in ranges])
    masses = rng.random(300) * t.sum()
    found = np.array([t.find_prefixsum_idx(float(m)) for m in masses])
    # the survey's hand example
    t8 = T(8)
    for i, v in enumerate([0.5, 1, 0.25, 2, 0, 3]):
        t8[i] = v
    np.savez(os.path.join(OUT, "replay_tree.npz"), cap=cap, idx=idx, val=val, value=leaves, ranges=ranges,
             sums=sums, masses=masses, found=found,
             t8=np.array([t8.sum(), t8.sum(0, 5), t8.find_prefixsum_idx(1.6)], dtype=np.float64))


def gen_buffer():
    PRB = R["replay_buffer"].PrioritizedReplayBuffer
    rng = np.random.default_rng(5)
    buf = quiet(PRB, 1000, 4.0, 10, 1e-7)
    nodes1 = rng.permutation(5000)[:300]
    pri1 = np.full(300, 2.0)
    buf.add_all(dict(zip(nodes1.tolist(), pri1.tolist())))
    leaves1 = np.array(buf._it_sum._value, dtype=np.float64)
    upd_nodes = nodes1[rng.permutation(300)[:120]]
    upd_pri = np.exp(rng.normal(0, 2, 120))          # spans the clip range
    buf.update_priorities(dict(zip(upd_nodes.tolist(), upd_pri.tolist())))
    leaves2 = np.array(buf._it_sum._value, dtype=np.float64)
    mm = np.array([buf.get_min_priority(), buf.get_max_priority(), buf._min_priority, buf._max_priority])
    nodes2 = np.arange(6000, 6100)
    p2 = buf.get_min_priority() + (buf.get_max_priority() - buf.get_min_priority()) * 0.95
    buf.add_all(dict(zip(nodes2.tolist(), [p2] * 100)))
    leaves3 = np.array(buf._it_sum._value, dtype=np.float64)
    # a proportional draw with the consumed `random` stream recorded
    rec_u, rec_i = [], []
    orig_random, orig_randint = random.random, random.randint
    random.seed(123)
    def rr():
        u = orig_random(); rec_u.append(u); return u
    def ri(a, b):
        v = orig_randint(a, b); rec_i.append(v); return v
    random.random, random.randint = rr, ri
    try:
        res = buf._sample_proportional(64)
    finally:
        random.random, random.randint = orig_random, orig_randint
    p_total = buf._it_sum.sum(0, len(buf._storage) - 1)
    np.savez(os.path.join(OUT, "replay_buffer.npz"), size=1000, alpha=4.0, max_p=10.0, min_p=1e-7,
             nodes1=nodes1, pri1=pri1, leaves1=leaves1, upd_nodes=upd_nodes, upd_pri=upd_pri,
             leaves2=leaves2, minmax=mm, nodes2=nodes2, p2=p2, leaves3=leaves3,
             draw_n=64, draw_uniforms=np.array(rec_u), draw_randints=np.array(rec_i, dtype=np.int64),
             draw_result=np.array(sorted(res), dtype=np.int64), p_total=p_total,
             storage=np.array(buf._storage, dtype=np.int64))


def make_edge_stream(V=400, E=3000, seed=1):
    """Power-law-ish stream relabelled to first-appearance order (reddit.py:101-113 precondition)."""
    rng = np.random.default_rng(seed)
    w = 1.0 / np.arange(1, V + 1) ** 0.8
    w /= w.sum()
    s = rng.choice(V, E, p=w)
    d = rng.choice(V, E, p=w)
    relabel = {}
    for a, b in zip(s.tolist(), d.tolist()):
        for x in (a, b):
            if x not in relabel:
                relabel[x] = len(relabel)
    s = np.array([relabel[x] for x in s.tolist()])
    d = np.array([relabel[x] for x in d.tolist()])
    return s, d, len(relabel)


def gen_edge_stream():
    DGE = R["dynamic_graph_edge"].DynamicGraphEdge
    s, d, V = make_edge_stream()
    snapshots = 25
    feats = torch.arange(V * 3, dtype=torch.float32).view(V, 3)
    targets = torch.arange(V, dtype=torch.int64).view(V, 1) % 5
    labelled = set(np.nonzero(np.arange(V) % 3 != 0)[0].tolist())
    g = quiet(DGE, snapshots, labelled)
    quiet(g.build, feats, targets, False, edge_timestamps=pd.DataFrame({"src": s, "dst": d}))
    n_nodes, n_edges, newv, feat_rows = [g.current_subgraph.n], [g.current_subgraph.number_of_edges()], [sorted(g.new_vertices)], []
    added = []
    for k in range(10):
        quiet(g.evolve)
        n_nodes.append(g.current_subgraph.n)
        n_edges.append(g.current_subgraph.number_of_edges())
        newv.append(sorted(g.new_vertices))
        v, lab = g.get_added_vertices(3)
        added.append((np.asarray(v), np.asarray(lab)))
    np.savez(os.path.join(OUT, "edge_stream.npz"), src=s, dst=d, V=V, snapshots=snapshots,
             labelled=np.array(sorted(labelled)),
             log_src=np.array(g.current_subgraph.src), log_dst=np.array(g.current_subgraph.dst),
             n_nodes=np.array(n_nodes), n_edges=np.array(n_edges),
             newv_flat=np.concatenate([np.array(x, dtype=np.int64) for x in newv]),
             newv_len=np.array([len(x) for x in newv]),
             feat=g.current_subgraph.ndata["feat"].numpy(), target=g.current_subgraph.ndata["target"].numpy(),
             added3_vertices=added[-1][0], added3_labelled=added[-1][1])


def gen_vertex_stream():
    DGV = R["dynamic_graph_vertex"].DynamicGraphVertex
    rng = np.random.default_rng(9)
    V = 203
    parent = R["RecGraph"](n=V)
    ts = {int(v): float(t) for v, t in zip(range(V), rng.integers(0, 40, V))}   # ties -> stable sort matters
    labelled = set(np.nonzero(rng.random(V) < 0.6)[0].tolist())
    g = quiet(DGV, parent, 10, labelled)
    quiet(g.build, vertex_timestamps=ts)
    chunks = [np.array(c, dtype=np.int64) for c in g.snapshot_vertices]
    act, s2o, o2s_probe = [len(g.evolving_vertices)], [], []
    for k in range(4):
        quiet(g.evolve)
        act.append(len(g.evolving_vertices))
    s2o = np.asarray(g.get_subgraph_to_original_map())
    probe = s2o[::7]
    o2s = np.asarray(g.get_original_to_subgraph_map()[probe])
    v, lab = g.get_added_vertices(2)
    order = np.argsort(v)
    np.savez(os.path.join(OUT, "vertex_stream.npz"), V=V, snapshots=10,
             ts_vertex=np.array(list(ts.keys())), ts_time=np.array(list(ts.values())),
             labelled=np.array(sorted(labelled)), n_chunks=len(chunks),
             chunk_flat=np.concatenate(chunks), chunk_len=np.array([len(c) for c in chunks]),
             n_active=np.array(act), s2o=s2o, probe=probe, o2s=o2s,
             added2_vertices=np.asarray(v)[order], added2_labelled=np.asarray(lab)[order],
             len_graph=len(g))


def gen_train_test():
    TT = R["train_test_graph"]
    DGE = R["dynamic_graph_edge"].DynamicGraphEdge
    TT.SIZE_BUFFER = 4096
    s, d, V = make_edge_stream(V=600, E=6000, seed=3)
    snapshots = 20
    feats = torch.zeros(V, 2)
    targets = torch.zeros(V, 1, dtype=torch.int64)
    labelled = set(np.nonzero(np.arange(V) % 4 != 1)[0].tolist())
    np.random.seed(1)
    random.seed(1)
    g = quiet(DGE, snapshots, labelled)
    quiet(g.build, feats, targets, False, edge_timestamps=pd.DataFrame({"src": s, "dst": d}))
    tt = quiet(TT.TrainTestGraph, g, split=0.15, start_prior_alpha=4, end_prior_alpha=50, scale=1, max_priority=10)
    rec = dict(train_len=[], test_len=[], rbr=[], pbr=[], newn=[], alpha=[], minmax=[])
    for k in range(8):
        rec["train_len"].append(len(tt.get_train_set()))
        rec["test_len"].append(len(tt.get_test_set()))
        rec["rbr"].append(list(tt.draw_random_train_nodes(16)))
        rec["pbr"].append(list(tt.draw_priority_train_nodes(16)))
        rec["newn"].append(list(tt.get_new_train_nodes(5)))
        # a partial priority update, as the PBR trainer does per batch (pytorch/model.py:203-206)
        upd = {int(v): 0.1 + 0.01 * (int(v) % 37) for v in rec["pbr"][-1]}
        quiet(tt.update_priorities, upd)
        rec["alpha"].append(tt.prior_alpha)
        rec["minmax"].append([tt.priority_replay_buffer.get_min_priority(), tt.priority_replay_buffer.get_max_priority()])
        quiet(tt.evolve)
    final_train = np.array(tt.get_train_set(), dtype=np.int64)
    final_test = np.array(tt.get_test_set(), dtype=np.int64)
    pri = np.array(tt.dump_priorities(final_train.tolist()))
    pad = lambda L, n: np.array([list(x) + [-1] * (n - len(x)) for x in L], dtype=np.int64)
    np.savez(os.path.join(OUT, "train_test.npz"), src=s, dst=d, V=V, snapshots=snapshots,
             labelled=np.array(sorted(labelled)), size_buffer=4096,
             train_len=np.array(rec["train_len"]), test_len=np.array(rec["test_len"]),
             rbr=pad(rec["rbr"], 16), pbr=pad(rec["pbr"], 16), newn=pad(rec["newn"], 5),
             alpha=np.array(rec["alpha"]), minmax=np.array(rec["minmax"], dtype=np.float64),
             final_train=final_train, final_test=final_test, final_priorities=pri)


def gen_priority_strategies():
    """TrendPriority / HybridPriority of the reference (generate_priority.py:11-58).  They use the removed `np.float` alias,
    so it is restored for the duration of the call (numpy 2 in this container); nothing else is touched."""
    import numpy
    had = hasattr(numpy, "float")
    if not had:
        numpy.float = float
    try:
        from prioritized_replay import generate_priority as gp
        rng = np.random.default_rng(11)
        V = 60
        trend, hybrid = gp.TrendPriority(V, alpha=0.85), gp.HybridPriority(V, alpha=0.7, loss_contrib=0.4)
        nodes, losses, out_t, out_h = [], [], [], []
        for step in range(12):
            b = rng.permutation(V)[:16]
            l = (rng.random(16) * 3).astype(np.float64)
            nodes.append(b); losses.append(l)
            out_t.append(np.array(trend.get_priorities(b.tolist(), l.copy()), dtype=np.float64))
            out_h.append(np.array(hybrid.get_priorities(b.tolist(), l.copy()), dtype=np.float64))
        np.savez(os.path.join(OUT, "priority_strategies.npz"), V=V, nodes=np.array(nodes), losses=np.array(losses),
                 trend=np.array(out_t), hybrid=np.array(out_h), trend_avg=trend.avg, trend_values=trend.values)
    finally:
        if not had:
            del numpy.float


if __name__ == "__main__":
    gen_tree(); gen_buffer(); gen_edge_stream(); gen_vertex_stream(); gen_train_test(); gen_priority_strategies()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
