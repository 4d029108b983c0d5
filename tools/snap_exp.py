"""µs per snapshot insert (11,461 stream edges, symmetrised) into the live Reddit-shaped CSR: fused kernel vs general path."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from ogl_b200 import native

w = bench.WORKLOADS["reddit"]
src, dst = bench.gen_edges(w, "cuda")
for fused in ("1", "0"):
    os.environ["OGL_INSERT_FUSED"] = fused
    g = native.Graph(w["V"], 2 * w["E"] + (1 << 21))
    g.insert_vertices(w["V"])
    for a in range(0, w["E"], 1 << 21):
        g.insert_edges(src[a:a + (1 << 21)], dst[a:a + (1 << 21)], symmetric=True)
    torch.cuda.synchronize()
    r = bench.aux_snapshot_insert(g, w["V"])
    print("fused" if fused == "1" else "general", {k: round(v, 1) for k, v in r.items()})
    del g
