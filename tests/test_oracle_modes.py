"""The arithmetic modes as the oracle models them (CPU only): what rounding every stored GEMM operand to bf16 / TF32 / fp16 costs against
the unquantised fp64 evaluation of the same GraphSAGE-pool step.  This is the reasoning behind the benchmarked fp16 mode, kept as a test:
fp16 storage has TF32's ten explicit mantissa bits, so with the product's static loss scaling (csrc/plan.cu: grad_scale_for,
oracle/sage.py: grad_scale_for) it is as accurate as TF32 -- and WITHOUT the scaling small gradients fall into fp16's subnormal range
and lose accuracy.  (The device kernels are compared with these oracle modes, and with the unquantised oracle at the benchmarked
shape, in the -m gpu tests.)"""
import numpy as np
import torch

from oracle import sage as osage


def _problem(seed=0, F=96, H=80, C=11, B=64, f1=6, f2=4, V=900):
    rng = np.random.default_rng(seed)

    def hop(dst, f):
        es = rng.integers(0, V, (len(dst), f))
        nodes, pos = list(dst), {int(v): i for i, v in enumerate(dst)}
        lid = np.empty_like(es)
        for i in range(es.shape[0]):
            for j in range(f):
                v = int(es[i, j])
                if v not in pos:
                    pos[v] = len(nodes)
                    nodes.append(v)
                lid[i, j] = pos[v]
        return np.array(nodes), lid

    seeds = rng.choice(V, B, replace=False)
    n1, l1 = hop(seeds, f1)
    n0, l0 = hop(n1, f2)
    blocks = [dict(n_dst=len(n1), edge_src=torch.from_numpy(l0.reshape(-1)).long(), fanout=f2),
              dict(n_dst=B, edge_src=torch.from_numpy(l1.reshape(-1)).long(), fanout=f1)]
    feats = torch.from_numpy(rng.standard_normal((V, F)))
    labels = torch.from_numpy(rng.integers(0, C, B))
    params = osage.xavier_params(F, H, C, 1, seed, dtype=torch.float64)
    return params, feats[torch.from_numpy(n0)], blocks, labels


def _errors(quant, params, x, blocks, labels, ref, loss_factor=1.0):
    if loss_factor == 1.0:
        _, _, logits, grads, _ = osage.loss_and_grads(params, x, blocks, labels, quant=quant, dtype=torch.float64)
    else:
        # the same step with the loss scaled down (gradients `loss_factor` times smaller, as with a large global batch)
        p = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        n_eff = int(round(labels.numel() / loss_factor))
        osage._GSCALE[0] = osage.grad_scale_for(1.0 / n_eff) if quant == "fp16" else 1.0
        logits, _ = osage.forward(p, x, blocks, quant=quant)
        (osage.xent(logits, labels, "none", quant=quant).sum() / n_eff).backward()
        logits, grads = logits.detach(), {k: v.grad * (n_eff / labels.numel()) for k, v in p.items()}
    e_log = ((logits - ref[2]).abs().max() / ref[2].abs().max()).item()
    e_grad = max(((grads[k] - ref[3][k]).norm() / ref[3][k].norm()).item() for k in ("layers.1.fc_self.weight", "layers.1.fc_neigh.weight"))
    return e_log, e_grad


def test_fp16_storage_is_as_accurate_as_tf32_and_bf16_is_not():
    params, x, blocks, labels = _problem()
    ref = osage.loss_and_grads(params, x, blocks, labels, quant=None, dtype=torch.float64)
    e = {q: _errors(q, params, x, blocks, labels, ref) for q in ("tf32", "fp16", "bf16")}
    assert e["fp16"][0] <= 1.25 * e["tf32"][0] + 1e-6 and e["fp16"][1] <= 1.25 * e["tf32"][1] + 1e-6, e
    assert e["tf32"][0] < 1e-3 and e["fp16"][0] < 1e-3, e                      # north_star's rtol for the tensor-core modes
    assert e["bf16"][0] > 4 * e["tf32"][0] and e["bf16"][1] > 4 * e["tf32"][1], e


def test_loss_scaling_keeps_small_gradients_out_of_the_subnormal_range():
    params, x, blocks, labels = _problem(seed=1)
    ref = osage.loss_and_grads(params, x, blocks, labels, quant=None, dtype=torch.float64)
    factor = 1.0 / 8192                                                       # gradients 8192 times smaller: dlogits of a few 1e-6
    scaled = _errors("fp16", params, x, blocks, labels, ref, loss_factor=factor)
    tf32 = _errors("tf32", params, x, blocks, labels, ref, loss_factor=factor)
    assert scaled[1] <= 1.25 * tf32[1] + 1e-6, (scaled, tf32)
    # the same gradients rounded to fp16 WITHOUT the scale: the layer-1 weight gradients lose accuracy
    keep = osage.grad_scale_for
    osage.grad_scale_for = lambda a: 1.0
    try:
        unscaled = _errors("fp16", params, x, blocks, labels, ref, loss_factor=factor)
    finally:
        osage.grad_scale_for = keep
    assert unscaled[1] > 1.5 * scaled[1], (unscaled, scaled)
