"""Parity metrics of one device train step against the UNQUANTISED oracle (oracle/sage.py in fp64 with quant=None: the
reference's fp32 arithmetic, train/utils.py:63-64 + train/graphsage/pytorch/model.py:96-107, restated) -- test infrastructure.

Used by tests/test_gpu_parity_reddit.py and by bench.py's parity leg (untimed, after the measurement): both run the device step
through the C ABI on a sampled minibatch, read the sampled blocks back and hand them to the oracle, so the comparison is on the
same inputs.  What is reported, per tensor:

  max_err_of_scale   max |a - b| / max |b|                 (the `atol = rtol * scale` reading of north_star's rtol)
  rel_fro            ||a - b||_F / ||b||_F
  frac_outside       share of elements with |a - b| > rtol * |b| + rtol * max |b|   (rtol = 1e-3)

and for the gradients twice.  GraphSAGE-pool is piecewise linear: which neighbour wins a max-pool and which ReLUs are on select
a linear region, and the gradient JUMPS between regions.  A pre-activation that is zero within rounding error, or two
neighbours that tie within rounding error, put the device and the oracle in neighbouring regions -- a whole different input
row for that (vertex, feature) pair, not a rounding error; any two fp32 implementations (the reference on CPU and on GPU, say)
differ the same way.  So:

  grad_free     the oracle picks its own regions (includes the jumps; `argmax_flip_frac` / `relu_flip_frac` say how many)
  grad_pinned   the oracle is evaluated in the DEVICE's region (its max-pool slots and ReLU on/off patterns), after checking that
                every device choice is a valid one within the mode's rounding (`*_not_within_rounding` must be 0): what is left
                is the arithmetic error proper
"""
import numpy as np
import torch

from . import sage as osage

NAMES = ("fc_pool.weight", "fc_pool.bias", "fc_self.weight", "fc_self.bias", "fc_neigh.weight", "fc_neigh.bias")
RTOL = 1e-3


def tensor_err(a, b, rtol=RTOL):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = b.abs().max().item() if b.numel() else 0.0
    err = (a - b).abs()
    fro = (err.pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp(min=1e-300)).item()
    return {"max_err_of_scale": (err.max().item() / scale) if scale > 0 else 0.0, "rel_fro": fro,
            "frac_outside": (err > rtol * b.abs() + rtol * scale).double().mean().item(), "scale": scale}


def dict_from_flat(flat, dims):
    out, off = {}, 0
    for i in range(len(dims) - 1):
        fi, fo = dims[i], dims[i + 1]
        for n, shape in zip(NAMES, ((fi, fi), (fi,), (fo, fi), (fo,), (fo, fi), (fo,))):
            k = int(np.prod(shape))
            out["layers.%d.%s" % (i, n)] = flat[off:off + k].view(shape)
            off += k
    assert off == flat.numel()
    return out


def device_blocks(plan):
    """the minibatch the plan last sampled, in the oracle's block format (input layer first)"""
    blocks = []
    for hop in reversed(range(plan.L)):
        lid, _, _, f = plan.block_edges(hop)
        blocks.append(dict(n_dst=plan.level_nodes(hop).numel(), edge_src=lid.long().cpu(), fanout=f))
    return plan.level_nodes(plan.L).long().cpu(), blocks


def compare_step(plan, params, feats_cpu, labels_cpu, seeds_cpu, logits_dev, per_dev, grad_dev, slot_tol):
    """plan: a native.Plan that has just run sample + forward + loss_backward(1 / n_seeds) on `seeds_cpu`; params: the oracle-format
    parameter dict (fp32 CPU tensors) the plan is bound to; feats_cpu [V, F] fp32, labels_cpu [V] int64: what the feature store
    holds.  slot_tol: relative rounding of a stored hp value in the plan's mode (2^-8 bf16, 2^-10 tf32, 2^-20 fp32).
    Returns the metrics dict described in the module docstring."""
    dims, L = plan.dims, plan.L
    input_nodes, blocks = device_blocks(plan)
    x_in = feats_cpu[input_nodes]
    labels = labels_cpu[torch.as_tensor(seeds_cpu)]
    # pass 1: free-running fp64 oracle
    _, per_ref, logits_ref, grads_free, inter = osage.loss_and_grads(params, x_in, blocks, labels, quant=None, dtype=torch.float64)
    out = {"oracle": "oracle/sage.py, fp64, quant=None (the reference's fp32 path restated), same sampled blocks",
           "rtol": RTOL, "logits": tensor_err(logits_dev, logits_ref), "per_vertex_loss": tensor_err(per_dev, per_ref),
           "level_counts": [plan.level_nodes(lv).numel() for lv in range(L + 1)]}
    # the device's argmax slots: each must attain the oracle's maximum within the mode's rounding; the device's ReLU patterns: a
    # sign that differs from the oracle's must belong to a pre-activation that is zero within the mode's rounding
    flips, not_max, relu_flips, relu_bad = [], [], [], []
    for l in range(L):
        h = L - 1 - l
        n_dst = plan.level_nodes(h).numel()
        n_src = plan.level_nodes(h + 1).numel()
        for name, pre, rows in (("hp", inter[l]["hp_pre"], n_src),) + ((("out", inter[l]["out_pre"], n_dst),) if l < L - 1 else ()):
            cols = dims[l] if name == "hp" else dims[l + 1]
            on = plan.tensor("%s%d" % (name, l), rows=rows)[:, :cols].float().cpu() > 0
            differ = on != (pre > 0)
            relu_flips.append(differ.double().mean().item())
            relu_bad.append(int((differ & (pre.abs() > 4 * slot_tol * pre.abs().max())).sum()))
            blocks[l]["mask_" + name] = on
        arg = plan.tensor("arg%d" % l, rows=n_dst)[:, :dims[l]].long().cpu()
        arg[arg == 255] = -1
        ref_arg = inter[l]["arg"]
        assert torch.equal(arg < 0, ref_arg < 0), "layer %d: rows without in-edges differ" % l
        es = blocks[l]["edge_src"].view(n_dst, blocks[l]["fanout"])
        src_row = torch.gather(es, 1, arg.clamp(min=0))
        assert bool((src_row[arg >= 0] >= 0).all()), "layer %d: a device argmax slot is empty" % l
        hp_o = inter[l]["hp"].detach()
        picked = hp_o[src_row.clamp(min=0), torch.arange(dims[l])[None, :].expand_as(src_row)]
        mx = inter[l]["neigh"].detach()
        ok = (picked >= mx - 2 * slot_tol * mx.abs() - slot_tol * hp_o.abs().max()) | (arg < 0)
        not_max.append(int((~ok).sum()))
        live = arg >= 0
        # a slot differs "really" only if it points at a different source row (the same neighbour may be sampled twice)
        ref_row = torch.gather(es, 1, ref_arg.clamp(min=0))
        flips.append(((src_row != ref_row) & live).double().sum().item() / max(1.0, live.double().sum().item()))
        blocks[l]["arg"] = arg
    out["argmax_flip_frac"] = flips
    out["argmax_not_a_max_within_rounding"] = not_max
    out["relu_flip_frac"] = relu_flips                       # [hp0, out0, hp1, ...]
    out["relu_sign_not_within_rounding"] = relu_bad
    # pass 2: the oracle evaluated in the device's linear region
    _, _, _, grads_routed, _ = osage.loss_and_grads(params, x_in, blocks, labels, quant=None, dtype=torch.float64)
    got = dict_from_flat(grad_dev.detach().double().cpu(), dims)
    gf, gr = {}, {}
    for k in got:
        gf[k] = tensor_err(got[k], grads_free[k])
        gr[k] = tensor_err(got[k], grads_routed[k])
    out["grad_free"] = gf
    out["grad_pinned"] = gr
    out["grad_rel_fro_free_max"] = max(v["rel_fro"] for v in gf.values())
    out["grad_rel_fro_pinned_max"] = max(v["rel_fro"] for v in gr.values())
    out["grad_max_err_of_scale_pinned_max"] = max(v["max_err_of_scale"] for v in gr.values())
    out["grad_frac_outside_pinned_max"] = max(v["frac_outside"] for v in gr.values())
    return out


def summary(m):
    """the few numbers bench.py prints in its JSON line"""
    return {"oracle": m["oracle"], "rtol": m["rtol"], "level_counts": m["level_counts"],
            "logits_max_err_of_scale": m["logits"]["max_err_of_scale"], "logits_rel_fro": m["logits"]["rel_fro"],
            "logits_frac_outside_rtol": m["logits"]["frac_outside"],
            "loss_max_err_of_scale": m["per_vertex_loss"]["max_err_of_scale"],
            "grad_rel_fro_free": {k: v["rel_fro"] for k, v in m["grad_free"].items()},
            "grad_rel_fro_pinned": {k: v["rel_fro"] for k, v in m["grad_pinned"].items()},
            "grad_rel_fro_free_max": m["grad_rel_fro_free_max"], "grad_rel_fro_pinned_max": m["grad_rel_fro_pinned_max"],
            "grad_max_err_of_scale_pinned_max": m["grad_max_err_of_scale_pinned_max"],
            "grad_frac_outside_rtol_pinned_max": m["grad_frac_outside_pinned_max"],
            "argmax_flip_frac": m["argmax_flip_frac"], "argmax_not_a_max_within_rounding": m["argmax_not_a_max_within_rounding"],
            "relu_flip_frac": m["relu_flip_frac"], "relu_sign_not_within_rounding": m["relu_sign_not_within_rounding"]}
