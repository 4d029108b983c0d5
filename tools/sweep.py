"""BASELINE.json configs[4]: synthetic power-law edge-stream sweep -- streaming-CSR insert and uniform k-neighbour sample
throughput against the HBM roofline, on one GPU.

    python tools/sweep.py [--edges 10000000 100000000 1000000000] [--batch 2097152]

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


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--edges", type=int, nargs="+", default=[10_000_000, 100_000_000, 1_000_000_000])
    ap.add_argument("--batch", type=int, default=1 << 21)
    a = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    for E in a.edges:
        print(json.dumps(run(E, a.batch, peak)), flush=True)
        torch.cuda.empty_cache()
