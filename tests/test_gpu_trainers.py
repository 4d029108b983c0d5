"""The drop-in surface end to end on the GPU: utils.init() -> GraphSAGE + the four trainer policies over a
streaming edge graph and a streaming vertex graph, driven exactly like the reference's snapshot loop
(train/__main__.py:161-196)."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _planted(V, E, F, C, seed):
    """labels = communities; edges mostly intra-community; features = noisy one-hot of the label"""
    rng = np.random.default_rng(seed)
    y = rng.integers(0, C, V)
    src = rng.integers(0, V, E)
    same = rng.random(E) < 0.85
    cand = rng.integers(0, V, (E, 8))
    pick = np.argmax(y[cand] == y[src][:, None], axis=1)
    dst = np.where(same, cand[np.arange(E), pick], rng.integers(0, V, E))
    x = rng.standard_normal((V, F)).astype(np.float32) * 0.7
    x[np.arange(V), y % F] += 1.5
    return src.astype(np.int64), dst.astype(np.int64), x, y.reshape(-1, 1).astype(np.int64)


def _relabel_first_appearance(src, dst, x, y):
    order = {}
    for a, b in zip(src.tolist(), dst.tolist()):
        for v in (a, b):            # the reference's np.unique per slice sorts ids; dense-in-order is the precondition
            if v not in order:
                order[v] = len(order)
    perm = np.array(sorted(order, key=order.get), dtype=np.int64)
    m = np.full(x.shape[0], -1, dtype=np.int64)
    m[perm] = np.arange(len(perm))
    return m[src], m[dst], x[perm], y[perm]


@pytest.mark.parametrize("faithful", [True, False])
def test_snapshot_loop_four_policies_edge_stream(tmp_path, faithful):
    import ogl_b200
    from ogl_b200 import config
    from ogl_b200.graph import train_test_graph as ttg
    config.set_faithful(faithful)
    config.set_precision("bf16")
    old = ttg.SIZE_BUFFER
    ttg.SIZE_BUFFER = 1 << 14
    try:
        random.seed(1)
        np.random.seed(1)
        torch.manual_seed(1)
        V, E, F, C, H = 1200, 9000, 16, 4, 32
        src, dst, x, y = _planted(V, E, F, C, seed=1)
        snapshots = 12
        src, dst, x, y = _relabel_first_appearance(src, dst, x, y)      # precondition of the reference (reddit.py:101-113)
        Vn = len(x)
        labelled = set(range(Vn))
        GraphSAGE, RandomT, PrioT, NoRehT, FullT, act = ogl_b200.init(ogl_b200.Lib_supported.PYTORCH, True, -1)
        dyn = ogl_b200.DynamicGraphEdge(snapshots, labelled)
        dyn.build(x, y, edge_timestamps={"src": src, "dst": dst})
        gu = ttg.TrainTestGraph(dyn, split=0.15, start_prior_alpha=4, end_prior_alpha=50, scale=1, max_priority=10)
        mk = lambda: GraphSAGE(F, H, C, 1, act, 0, "pool").cuda()
        kw = dict(cuda=True, batch_full=256, n_workers=0)
        trainers = [RandomT(mk(), 20, 32, y, 5, **kw),
                    PrioT(mk(), 20, 32, y, 5, ogl_b200.LossPriority(), full_pass=2, **kw),
                    NoRehT(mk(), 20, 32, y, 5, **kw),
                    FullT(mk(), 2, 32, y, 5, **kw)]
        names = [t.get_model() for t in trainers]
        assert names == ["random", "prioritized", "no_rehersal", "offline"]
        for t in trainers:
            t.build_optimizer()
        out = str(tmp_path / "res.csv")
        f1 = {}
        for snap in range(snapshots - 1):
            for t in trainers:
                t.train_timestep(gu)
                assert t.delay > 0
                r = t.evaluate(gu, out)
                if r is not None:
                    f1.setdefault(t.get_model(), []).append(r)
            gu.evolve()
        rows = open(out).read().strip().split("\n")
        assert len(rows) >= len(trainers) * (snapshots - 2)
        assert rows[0].split(";")[0] == "random" and len(rows[0].split(";")) == 4
        # the planted structure is learnable: rehearsal policies end well above chance (1/C = 0.25 macro-F1)
        for name in ("random", "prioritized", "offline"):
            assert np.mean(f1[name][-3:]) > 0.5, (name, f1[name])
        for t in trainers:
            for p in t.graphsage_model.parameters():
                assert torch.isfinite(p).all()
    finally:
        ttg.SIZE_BUFFER = old
        config.set_faithful(True)


def test_vertex_stream_training_and_next_snapshot_eval(tmp_path):
    import ogl_b200
    from ogl_b200 import config
    from ogl_b200.graph import train_test_graph as ttg
    config.set_faithful(True)
    old = ttg.SIZE_BUFFER
    ttg.SIZE_BUFFER = 1 << 12
    try:
        random.seed(2)
        np.random.seed(2)
        torch.manual_seed(2)
        V, E, F, C, H = 900, 7000, 12, 3, 16
        src, dst, x, y = _planted(V, E, F, C, seed=7)
        y[::5] = -1                                             # unlabelled vertices
        labelled = set(np.nonzero(y.reshape(-1) >= 0)[0].tolist())
        pg = ogl_b200.ParentGraph.from_undirected(src, dst, V)
        pg.ndata["feat"], pg.ndata["target"] = x, y
        ts = {int(v): float(t) for v, t in enumerate(np.random.default_rng(0).permutation(V))}
        GraphSAGE, RandomT, PrioT, NoRehT, FullT, act = ogl_b200.init(ogl_b200.Lib_supported.PYTORCH, True, -1)
        dyn = ogl_b200.DynamicGraphVertex(pg, 10, labelled)
        dyn.build(vertex_timestamps=ts)
        dyn_test = ogl_b200.DynamicGraphVertex(pg, 10, labelled)
        dyn_test.build(vertex_timestamps=ts)
        delta = 2
        for _ in range(delta):
            dyn_test.evolve()
        gu = ttg.TrainTestGraph(dyn, split=0.15, start_prior_alpha=4, end_prior_alpha=50, scale=1, max_priority=10)
        t = RandomT(GraphSAGE(F, H, C, 1, act, 0, "pool").cuda(), 25, 32, y, 6, cuda=True, batch_full=128, n_workers=0)
        t.build_optimizer()
        out = str(tmp_path / "v.csv")
        scores = []
        for snap in range(6):
            t.train_timestep(gu)
            scores.append(t.evaluate(gu, out))
            t.evaluate_next_snapshots(dyn_test, delta, out)
            gu.evolve()
            dyn_test.evolve()
        assert scores[-1] is not None and scores[-1] > 0.5
        assert os.path.getsize(out) > 0
    finally:
        ttg.SIZE_BUFFER = old


def test_train_step_honours_the_blocks_it_is_given():
    """the reference's train_step(graph, blocks, input_nodes, seeds, ...) trains on the minibatch the loader produced
    (pytorch/model.py:77-107): blocks sampled by the trainer's own plan are trained as given, stale / foreign blocks are resampled"""
    import ogl_b200
    from ogl_b200.sampling import MultiLayerNeighborSampler, NodeDataLoader
    ogl_b200.config.set_precision("fp32")
    try:
        V, E, F, C, H = 700, 6000, 10, 3, 12
        src, dst, x, y = _planted(V, E, F, C, seed=3)
        dg = ogl_b200.DeviceGraph(V, 2 * E, F)
        dg.add_nodes(V, {"feat": x, "target": y})
        dg.add_edges(src, dst, symmetric=True)
        GraphSAGE, RandomT, _, _, _, act = ogl_b200.init(ogl_b200.Lib_supported.PYTORCH, True, -1)
        torch.manual_seed(1)
        model = GraphSAGE(F, H, C, 1, act, 0, "pool").cuda()
        t = RandomT(model, 1, 32, y, 5, cuda=True, batch_full=64, n_workers=0)
        t.build_optimizer()
        plan = t._train_plan(dg)
        loader = NodeDataLoader(dg, np.arange(64), MultiLayerNeighborSampler([5, 5]), batch_size=32, plan=plan)
        it = iter(loader)
        input_nodes, seeds, blocks = next(it)
        before = model._flat.clone()
        frontier = plan.level_nodes(2).clone()
        # the oracle's loss gradient on exactly these blocks = the direction Adam's first step takes
        assert t.train_step(dg, blocks, input_nodes, seeds, None) == "trained on the given blocks"
        assert torch.equal(plan.level_nodes(2), frontier), "the given minibatch was resampled"
        step1 = model._flat - before
        assert float(step1.abs().max()) > 0
        from oracle import sage as osage
        params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        ob = [dict(n_dst=b.number_of_dst_nodes(), edge_src=b._lid.long().cpu(), fanout=b.fanout) for b in blocks]
        p0 = {k: before[o:o + v.numel()].view(v.shape).cpu() for (k, v), o in zip(params.items(), np.cumsum([0] + [v.numel() for v in params.values()])[:-1])}
        _, _, _, gref, _ = osage.loss_and_grads(p0, torch.as_tensor(x)[input_nodes.cpu()], ob, torch.as_tensor(y)[seeds.cpu()].flatten(), dtype=torch.float64)
        gflat = torch.cat([gref[k].reshape(-1) for k in params])
        big = gflat.abs() > 1e-3 * gflat.abs().max()
        assert bool((torch.sign(step1.cpu().double()[big]) == -torch.sign(gflat[big])).all()), "first Adam step is not along -grad of the given blocks"
        # stale blocks (the plan has sampled another minibatch since) are not replayable
        next(it)
        assert t.train_step(dg, blocks, input_nodes, seeds, None) == "resampled"
    finally:
        ogl_b200.config.set_precision("tf32")
