"""N > 1 host logic on CPU (gloo, world_size 2): seed sharding, the gradient all-reduce contract (sum of per-rank
gradients scaled by 1 / global batch == mean gradient of the global batch) and the loss all-gather used by PBR."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "ogl_parallel", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "online-gnn-learning_b200", "parallel.py"))
    par = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(par)
    assert par.world() == world and par.rank() == rank
    # 1. shards of a global batch are disjoint, ordered, and cover it (uneven size on purpose)
    glob = np.arange(1000, 1000 + 37, dtype=np.int64)
    mine = par.shard(glob)
    # 2. gradient contract: per-sample gradients g_i, local "backward" produces sum_i g_i / global_batch
    rng = np.random.default_rng(7)
    per_sample = torch.from_numpy(rng.standard_normal((37, 11)))
    local = per_sample[mine - 1000].sum(0) / 37.0
    total = par.allreduce_grads(local.clone())
    # 3. PBR: (vertex, loss) pairs gathered in rank order on every rank
    v, l = par.allgather_losses(torch.from_numpy(mine), torch.from_numpy(mine.astype(np.float64) * 0.5))
    out[rank] = (mine.tolist(), total.tolist(), v.tolist(), l.tolist(), per_sample.mean(0).tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_two_host_logic():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        r0, r1 = out[0], out[1]
    assert r0[0] + r1[0] == list(range(1000, 1037)) and abs(len(r0[0]) - len(r1[0])) <= 1
    assert np.allclose(r0[1], r0[4]) and np.allclose(r1[1], r0[4])           # all-reduced == mean gradient, on both ranks
    assert r0[2] == r1[2] == list(range(1000, 1037)) and r0[3] == r1[3]     # identical PBR update stream everywhere


def test_single_process_is_a_noop():
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ogl_parallel1", os.path.join(root, "online-gnn-learning_b200", "parallel.py"))
    par = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(par)
    g = torch.ones(5)
    assert par.world() == 1 and par.rank() == 0 and par.allreduce_grads(g) is g
    assert par.shard(list(range(10))) == list(range(10))
    assert par.shard(list(range(10)), 1, 3) == [4, 5, 6] and par.shard(list(range(10)), 0, 3) == [0, 1, 2, 3]
