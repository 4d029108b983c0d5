"""BASELINE.json configs[4]: synthetic power-law edge-stream sweep -- streaming-CSR insert and uniform k-neighbour sample
throughput against the HBM roofline, on one GPU or (--gpus N) on N GPUs with a destination-range-sharded CSR.

    python tools/sweep.py [--edges 10000000 100000000 1000000000] [--batch 2097152] [--gpus N]

For each stream size E (V = E / 16 vertices, endpoints ~ (rank + 50)^-0.83): the stream is generated batch by batch on the
GPU (untimed), every batch is inserted symmetrised (2 directed edges per stream edge) and timed with CUDA events; then
2^20 random rows x 25 picks are sampled (indices + edge ids).  One JSON line per size."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ogl_b200
from ogl_b200 import native


def run(E, batch, hbm_peak):
    V = max(E // 16, 1024)
    g = torch.Generator(device="cuda").manual_seed(1)
    wt = (torch.arange(V, device="cuda", dtype=torch.float64) + 50.0) ** -0.83
    cdf = torch.cumsum(wt / wt.sum(), 0).float()
    del wt
    perm = torch.randperm(V, generator=g, device="cuda")
    graph = native.Graph(V, 2 * E)
    graph.insert_vertices(V)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 0.0
    l0 = ogl_b200.kernel_launches()
    for a in range(0, E, batch):
        n = min(batch, E - a)
        src = perm[torch.searchsorted(cdf, torch.rand(n, generator=g, device="cuda")).clamp_(max=V - 1)]
        dst = perm[torch.searchsorted(cdf, torch.rand(n, generator=g, device="cuda")).clamp_(max=V - 1)]
        e0.record()
        graph.insert_edges(src, dst, symmetric=True)
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    launches = ogl_b200.kernel_launches() - l0
    assert graph.num_edges == 2 * E
    st = graph.stats()
    deg = graph.degrees()
    rows = torch.randint(0, V, (1 << 20,), generator=g, device="cuda", dtype=torch.int64)
    for _ in range(2):
        native.sample_neighbors(graph, rows, 25, seed=3, step=0, hop=0)
    torch.cuda.synchronize()
    e0.record()
    for i in range(5):
        native.sample_neighbors(graph, rows, 25, seed=3, step=i + 1, hop=0)
    e1.record()
    torch.cuda.synchronize()
    sms = e0.elapsed_time(e1) / 5
    picks = (1 << 20) * 25
    ins_alg = 2 * E * (16 + 8 + 8)                      # per directed edge: (src, dst) pair read, 8-byte entry write, degree RMW
    smp_alg = (1 << 20) * 20 + picks * 8 + picks * 12
    smp_sec = (1 << 20) * 64 + picks * 32 + picks * 12
    return {"stream_edges": E, "vertices": V, "batch": batch, "max_degree": int(deg.max()), "insert_ms": ms,
            "insert_stream_edges_per_s": E / (ms * 1e-3), "insert_algorithmic_gbs": ins_alg / (ms * 1e-3) / 1e9,
            "insert_frac_of_hbm_peak": ins_alg / (ms * 1e-3) / 1e9 / hbm_peak, "insert_kernel_launches": launches,
            "relocations": st["relocations"], "pool_rebuilds": st["compactions"],
            "sample_ms": sms, "sample_picks_per_s": picks / (sms * 1e-3), "sample_algorithmic_gbs": smp_alg / (sms * 1e-3) / 1e9,
            "sample_sector_gbs": smp_sec / (sms * 1e-3) / 1e9, "sample_sector_frac_of_hbm_peak": smp_sec / (sms * 1e-3) / 1e9 / hbm_peak}


def run_sharded(E, batch, hbm_peak, rank, world):
    """N GPUs (torchrun, one process per GPU): the CSR is sharded by DESTINATION range -- rank r owns the in-edge rows of vertices
    [r V / N, (r + 1) V / N) -- and every rank ingests its own 1 / N of each stream batch (edges arriving at N ingest points).  A
    stream edge (u, v) becomes the directed edges u -> v and v -> u; each goes to the owner of its destination: bucket by owner,
    one NCCL all-to-all of (source, local row) pairs per batch, then the owner appends what it received to its shard with the same
    insert kernels as the single-GPU path (edge ids are per shard, in arrival order).  Sampling: every rank samples 2^20 / N of its
    own rows (a sampler front end would route seeds to owners the same way).  Times are device times, max over the ranks."""
    import torch.distributed as dist
    V = max(E // 16, 1024)
    rows_per = (V + world - 1) // world
    lo = rank * rows_per
    v_local = max(0, min(V, lo + rows_per) - lo)
    gen = torch.Generator(device="cuda").manual_seed(1 + rank)
    wt = (torch.arange(V, device="cuda", dtype=torch.float64) + 50.0) ** -0.83
    cdf = torch.cumsum(wt / wt.sum(), 0).float()
    del wt
    perm = torch.randperm(V, generator=torch.Generator(device="cuda").manual_seed(1), device="cuda")     # same scatter on every rank
    graph = native.Graph(max(v_local, 1), int(2.6 * E / world) + (1 << 22))
    graph.insert_vertices(max(v_local, 1))
    graph.set_source_bound(V)                         # local rows, global source ids
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ms_route = ms_insert = 0.0
    per_rank_batch = batch // world
    warm = torch.zeros(world, dtype=torch.int64, device="cuda")          # communicator set-up outside the timed region
    dist.all_to_all_single(torch.empty_like(warm), warm)
    torch.cuda.synchronize()
    for a in range(0, E, batch):
        n = min(per_rank_batch, max(0, (min(batch, E - a) + world - 1) // world))
        u = perm[torch.searchsorted(cdf, torch.rand(n, generator=gen, device="cuda")).clamp_(max=V - 1)]
        v = perm[torch.searchsorted(cdf, torch.rand(n, generator=gen, device="cuda")).clamp_(max=V - 1)]
        dist.barrier()
        e0.record()
        # directed edges of this rank's slice, bucketed by the owner of the destination
        src = torch.cat([u, v])
        dst = torch.cat([v, u])
        owner = torch.div(dst, rows_per, rounding_mode="floor")
        order = torch.sort(owner, stable=True).indices
        send = torch.stack([src[order], dst[order] - owner[order] * rows_per], 1).contiguous()
        counts = torch.bincount(owner, minlength=world)
        recv_counts = torch.empty_like(counts)
        dist.all_to_all_single(recv_counts, counts)
        sc, rc = counts.tolist(), recv_counts.tolist()
        recv = torch.empty(sum(rc), 2, dtype=torch.int64, device="cuda")
        dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc)
        e1.record()
        if recv.shape[0]:
            graph.insert_edges(recv[:, 0].contiguous(), recv[:, 1].contiguous(), symmetric=False)
        e2.record()
        torch.cuda.synchronize()
        ms_route += e0.elapsed_time(e1)
        ms_insert += e1.elapsed_time(e2)
    local_edges = graph.num_edges
    rows = torch.randint(0, max(v_local, 1), ((1 << 20) // world,), generator=gen, device="cuda", dtype=torch.int64)
    for _ in range(3):
        native.sample_neighbors(graph, rows, 25, seed=3, step=0, hop=0)
    dist.barrier()
    torch.cuda.synchronize()
    e0.record()
    for i in range(5):
        native.sample_neighbors(graph, rows, 25, seed=3, step=i + 1, hop=0)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ms_route, ms_insert, e0.elapsed_time(e1) / 5], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot = torch.tensor([float(local_edges)], device="cuda", dtype=torch.float64)
    dist.all_reduce(tot)
    ms_route, ms_insert, sms = (float(x) for x in t)
    ms = ms_route + ms_insert
    picks = ((1 << 20) // world) * world * 25
    ins_alg = 2 * E * (16 + 8 + 8)
    smp_sec = ((1 << 20) // world) * world * 64 + picks * 32 + picks * 12
    return {"n_gpus": world, "sharding": "destination range, all-to-all edge routing (NCCL)", "stream_edges": E, "vertices": V, "batch": batch,
            "directed_edges_stored": int(tot.item()), "route_ms": ms_route, "insert_ms": ms_insert,
            "insert_stream_edges_per_s": E / (ms * 1e-3), "insert_only_stream_edges_per_s": E / (ms_insert * 1e-3),
            "insert_algorithmic_gbs": ins_alg / (ms * 1e-3) / 1e9, "insert_frac_of_aggregate_hbm_peak": ins_alg / (ms * 1e-3) / 1e9 / (hbm_peak * world),
            "sample_ms": sms, "sample_picks_per_s": picks / (sms * 1e-3), "sample_sector_gbs": smp_sec / (sms * 1e-3) / 1e9,
            "sample_sector_frac_of_aggregate_hbm_peak": smp_sec / (sms * 1e-3) / 1e9 / (hbm_peak * world)}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--edges", type=int, nargs="+", default=[10_000_000, 100_000_000, 1_000_000_000])
    ap.add_argument("--batch", type=int, default=1 << 21)
    ap.add_argument("--gpus", type=int, default=1, help="N > 1: run under torchrun --nproc-per-node N (destination-range sharded CSR)")
    a = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if a.gpus > 1 and world == 1:
        import subprocess
        sys.exit(subprocess.call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(a.gpus),
                                  "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.abspath(__file__)] + sys.argv[1:]))
    if world > 1:
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        for E in a.edges:
            r = run_sharded(E, a.batch, peak, dist.get_rank(), world)
            if dist.get_rank() == 0:
                print(json.dumps(r), flush=True)
            torch.cuda.empty_cache()
        dist.destroy_process_group()
        sys.exit(0)
    for E in a.edges:
        print(json.dumps(run(E, a.batch, peak)), flush=True)
        torch.cuda.empty_cache()
