"""Data-parallel training over the GPUs of one box: one process per GPU (torchrun), replicated
streaming graph + feature store + weights, every rank trains its own shard of the global target-vertex batch, the
flat fp32 gradient buffer is summed with ONE NCCL all-reduce (loss_scale = 1 / global batch makes the sum the mean
gradient) before the replicated fused Adam step.  PBR keeps its sum trees identical on all ranks by all-gathering the
(vertex id, loss) pairs.  The reference has no multi-GPU path (single process)."""
import torch
import torch.distributed as dist


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(seeds, r=None, w=None):
    """rank r's contiguous slice of a global seed list (equal sizes up to the remainder, order preserved)"""
    r = rank() if r is None else r
    w = world() if w is None else w
    n = len(seeds)
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return seeds[lo:lo + base + (1 if r < rem else 0)]


def allreduce_grads(flat_grad, group=None):
    """sum the flat gradient buffer over the ranks (in place); no-op on one rank"""
    if world() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return flat_grad


def allgather_losses(vertex_ids, losses, group=None):
    """every rank receives all (vertex id, per-vertex loss) pairs of the step, in rank order (PBR priority updates
    must be applied identically on every replica)"""
    if world() == 1:
        return vertex_ids, losses
    w = world()
    n = torch.tensor([vertex_ids.numel()], dtype=torch.int64, device=vertex_ids.device)
    sizes = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(sizes, n, group=group)
    m = int(max(int(s.item()) for s in sizes))
    pad_v = torch.full((m,), -1, dtype=vertex_ids.dtype, device=vertex_ids.device)
    pad_l = torch.zeros(m, dtype=losses.dtype, device=losses.device)
    pad_v[:vertex_ids.numel()] = vertex_ids
    pad_l[:losses.numel()] = losses
    vs = [torch.empty_like(pad_v) for _ in range(w)]
    ls = [torch.empty_like(pad_l) for _ in range(w)]
    dist.all_gather(vs, pad_v, group=group)
    dist.all_gather(ls, pad_l, group=group)
    keep = [int(s.item()) for s in sizes]
    return torch.cat([v[:k] for v, k in zip(vs, keep)]), torch.cat([l[:k] for l, k in zip(ls, keep)])


def train_step(plan, graph, features, local_seeds, global_batch, flat_grad, per_vertex_out=None, loss_sum_out=None):
    """one data-parallel step: local sample/forward/backward -> gradient all-reduce -> replicated Adam"""
    w = world()
    plan.train_step(graph, features, local_seeds, loss_scale=1.0 / float(global_batch), do_step=(w == 1),
                    per_vertex_out=per_vertex_out, loss_sum_out=loss_sum_out)
    if w > 1:
        allreduce_grads(flat_grad)
        plan.adam_step()
