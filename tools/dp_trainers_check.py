"""torchrun --nproc-per-node 2 tools/dp_trainers_check.py -- the four drop-in trainers in data-parallel mode on a planted-partition
edge stream: every rank must end every timestep with bit-identical weights and (PBR) identical priority sum trees; F1 above chance."""
import json
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import ogl_b200
    from ogl_b200 import config
    from ogl_b200.graph import train_test_graph as ttg
    from test_gpu_trainers import _planted, _relabel_first_appearance
    config.set_faithful(True)
    config.set_precision("tf32")
    ttg.SIZE_BUFFER = 1 << 12
    random.seed(1); np.random.seed(1); torch.manual_seed(1)
    V, E, F, C, H = 1200, 9000, 12, 3, 16
    src, dst, x, y = _planted(V, E, F, C, seed=5)
    src, dst, x, y = _relabel_first_appearance(src, dst, x, y)
    GraphSAGE, RandomT, PrioT, NoRehT, FullT, act = ogl_b200.init(ogl_b200.Lib_supported.PYTORCH, True, local)
    dyn = ogl_b200.DynamicGraphEdge(12, set(range(len(x))))
    dyn.build(x, y, edge_timestamps={"src": src, "dst": dst})
    gu = ttg.TrainTestGraph(dyn, split=0.15, start_prior_alpha=4, end_prior_alpha=50, scale=1, max_priority=10)
    mk = lambda: GraphSAGE(F, H, C, 1, act, 0, "pool").cuda()
    kw = dict(cuda=True, batch_full=128, n_workers=0)
    trainers = [RandomT(mk(), 25, 32, y, 5, **kw), PrioT(mk(), 25, 32, y, 5, ogl_b200.LossPriority(), full_pass=2, **kw),
                NoRehT(mk(), 25, 32, y, 5, **kw), FullT(mk(), 1, 32, y, 5, **kw)]
    for t in trainers:
        t.build_optimizer()
    ok, f1 = True, {}
    for step in range(10):
        for t in trainers:
            t.train_timestep(gu)
        for t in trainers:
            cs = t.graphsage_model._flat.view(torch.int32).to(torch.int64).sum().reshape(1)
            every = [torch.zeros_like(cs) for _ in range(dist.get_world_size())]
            dist.all_gather(every, cs)
            ok = ok and all(int(e) == int(every[0]) for e in every)
        leaves = gu.priority_replay_buffer._it_sum._t.values().sum().reshape(1)
        every = [torch.zeros_like(leaves) for _ in range(dist.get_world_size())]
        dist.all_gather(every, leaves)
        ok = ok and all(float(e) == float(every[0]) for e in every)
        for t in trainers:
            f1[t.get_model()] = t.evaluate(gu, None)
        if step + 1 < len(gu):
            gu.evolve()
    if dist.get_rank() == 0:
        print(json.dumps({"replicas_identical": bool(ok), "f1": f1, "world": dist.get_world_size()}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
