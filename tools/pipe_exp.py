"""Where does a train step's time go?  begin-only / finish-only / fused / pipelined replays on the Reddit-shaped workload
(python tools/pipe_exp.py [steps]); prints ms per step for each."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import ogl_b200
from ogl_b200 import native

K = int(sys.argv[1]) if len(sys.argv) > 1 else 100
w = bench.WORKLOADS["reddit"]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
src, dst = bench.gen_edges(w, dev)
V, E = w["V"], w["E"]
g = native.Graph(V, 2 * E)
g.insert_vertices(V)
for a in range(0, E, 1 << 21):
    g.insert_edges(src[a:a + (1 << 21)], dst[a:a + (1 << 21)], symmetric=True)
feats, labels = bench.gen_features(w, dev)
fs = native.Features(V, w["F"], ogl_b200.OGL_BF16)
fs.write(0, feats, labels)
del feats, src, dst
params = bench.init_params(w)
flat = torch.cat([params[f"layers.{i}.{n}"].reshape(-1).float() for i in range(2) for n in bench.NAMES]).to(dev)
grad = torch.zeros_like(flat)
plan = native.Plan([w["F"], w["H"], w["C"]], w["fanouts"], w["B"], V, mode=ogl_b200.OGL_BF16, seed=11)
plan.bind_params(flat, grad)
B = w["B"]
batches = [torch.as_tensor(b).to(dev) for b in bench.seed_batches(w, K + 3, 0, 1)]
loss = torch.zeros(1, device=dev)


def timed(fn, n=K):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def fused(i):
    plan.train_step(g, fs, batches[i], loss_scale=1.0 / B, do_step=True, loss_sum_out=loss)


def begin(i):
    plan.step_begin(g, fs, batches[i])


def finish(i):
    plan.step_finish(fs, 1.0 / B, do_step=True, loss_sum_out=loss)


pipe = ogl_b200.parallel.Pipeline(plan, g, fs, grad, B)


def piped(n):
    pipe.begin(batches[0])
    for i in range(n):
        pipe.finish(batches[i + 1] if i + 1 < n else None, loss_sum_out=loss)


for i in range(3):
    fused(i)
res = {}
res["fused"] = timed(fused)
res["begin_only"] = timed(begin)
plan.step_begin(g, fs, batches[0])
res["finish_only"] = timed(finish)
piped(3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); piped(K); e1.record(); torch.cuda.synchronize()
res["pipelined"] = e0.elapsed_time(e1) / K
for opt in ("side_stream",):
    plan.set_option(opt, 0)
    for i in range(3):
        fused(i)
    res["fused_no_" + opt] = timed(fused)
    plan.step_begin(g, fs, batches[0])
    res["finish_only_no_" + opt] = timed(finish)
    plan.set_option(opt, 1)
print({k: round(v, 4) for k, v in res.items()})
