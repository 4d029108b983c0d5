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


def make_peer_exchange(plan, flat_params, group=None):
    """NVLink peer-memory gradient exchange for `plan` (csrc/peer.cu): allocates the rank's peer-visible gradient buffer, binds it
    as the plan's gradient buffer, exchanges the CUDA IPC handles over the process group and maps the peers.  Returns the
    native.Peer (its `.grads` is the new flat gradient tensor).  One rank: a plain local buffer, no peers."""
    from . import _native as native
    w, r = world(), rank()
    peer = native.Peer(r, w, plan.n_params)
    if w > 1:
        mine = torch.frombuffer(bytearray(peer.handle()), dtype=torch.uint8).cuda()
        every = [torch.empty_like(mine) for _ in range(w)]
        dist.all_gather(every, mine, group=group)
        peer.connect(b"".join(bytes(t.cpu().numpy().tobytes()) for t in every))
        dist.barrier(group=group)
    plan.bind_params(flat_params, peer.grads)
    return peer


class Pipeline:
    """Train loop with everything that does not need fresh weights off the critical path (any number of GPUs).

    A step is split at the only point where it needs fresh weights: `begin` = sample + gather (reads the graph and the
    feature store only), `finish` = forward .. backward (.. Adam).  `begin` of step t+1 is a prefetch (Plan.prefetch: the
    plan's own stream and second buffer set), so it overlaps forward / backward of step t -- the job NodeDataLoader's worker
    processes do in the reference (pytorch/model.py:128-131).  With more than one rank the gradient exchange is bucketed:
    the all-reduce of everything but layer 0's fc_pool.weight overlaps the last weight-gradient GEMM, the small second
    bucket + Adam run on a communication stream; forward(t+1) waits for Adam(t).  Same arithmetic as the unpipelined loop
    (no stale gradients, same Philox step per minibatch)."""

    def __init__(self, plan, graph, features, flat_grad, global_batch, peer=None, force_split=False):
        """peer: a native.Peer from make_peer_exchange -> the gradient exchange + Adam is ONE kernel per bucket over NVLink peer
        memory (no NCCL on the data path); None -> NCCL all-reduce + ogl_plan_adam_step"""
        self.plan, self.graph, self.features, self.flat_grad = plan, graph, features, flat_grad
        self.peer = peer
        if peer is not None:
            assert flat_grad.data_ptr() == peer.grads.data_ptr(), "the plan must be bound to the peer group's gradient buffer"
        self.scale = 1.0 / float(global_batch)
        # force_split (experiments): run the data-parallel launch structure (head | bucket exchange | tail | bucket exchange) on ONE
        # rank with a one-rank peer group -- what the structure itself costs, without NVLink traffic or waiting for other ranks
        self.w = max(world(), 2) if force_split else world()
        self.comm = torch.cuda.Stream() if self.w > 1 else None
        import os
        self.split_launch = bool(int(os.environ.get("OGL_DP_SPLIT_LAUNCH", "0")))      # the older form: head | exchange | tail | exchange
        pieces = getattr(plan, "tail_pieces", [None])
        # measured on 2 and 8 B200 (tf32, Reddit shape): exchanging the last gradient in 256-row pieces is SLOWER (1.135 vs 1.014 ms
        # at 8 GPUs: three GEMM prologues, three reduces and three box-wide barriers cost more than the exposed exchange they hide)
        self.tail_pieces = len(pieces) if int(os.environ.get("OGL_DP_TAIL_PIECES", "0")) else 1
        self.ev_piece = [torch.cuda.Event() for _ in range(len(pieces))] if self.w > 1 else []
        self.ev_bwd = torch.cuda.Event()
        self.ev_head = torch.cuda.Event()
        self.ev_adam = torch.cuda.Event()
        self._begun = 0
        self._adam_pending = False

    def begin(self, seeds):
        self.plan.prefetch(self.graph, self.features, seeds)
        self._begun += 1

    def finish(self, next_seeds=None, per_vertex_out=None, loss_sum_out=None, scale=None):
        """scale: loss scale of THIS step (1 / its global batch) when it differs from the pipeline's (a ragged last minibatch)"""
        assert self._begun > 0, "Pipeline.finish() without begin()"
        if next_seeds is not None and self._begun < 2:
            self.begin(next_seeds)                       # enqueued first: overlaps this step's forward / backward
        self._begun -= 1
        main = torch.cuda.current_stream()
        keep_scale = self.scale
        if scale is not None:
            self.scale = float(scale)
        try:
            self._finish(main, per_vertex_out, loss_sum_out)
        finally:
            self.scale = keep_scale

    def _finish(self, main, per_vertex_out, loss_sum_out):
        if self.w == 1:
            self.plan.step_finish(self.features, self.scale, do_step=True, per_vertex_out=per_vertex_out, loss_sum_out=loss_sum_out)
            return
        if self.peer is not None and not self.split_launch and hasattr(self.plan, "step_finish_dp"):
            # one launch sequence (one CUDA graph) per step, the exchanges inside it (ogl_plan_step_finish_dp)
            self.plan.step_finish_dp(self.peer, self.features, self.scale, per_vertex_out=per_vertex_out, loss_sum_out=loss_sum_out)
            return
        if self._adam_pending:
            main.wait_event(self.ev_adam)                # weights (and the gradient buffer) of the previous step are settled
        # two gradient buckets: everything except layer 0's fc_pool.weight is final before the last weight-gradient GEMM
        # runs, so its all-reduce overlaps that GEMM; the small second bucket + Adam overlap the next step's start
        n0, n = self.plan.tail_params, self.plan.n_params
        if self.peer is not None:
            self.peer.wait_readers()                     # every peer has read this rank's gradients of the previous step
        self.plan.step_finish_head(self.features, self.scale, per_vertex_out=per_vertex_out, loss_sum_out=loss_sum_out)
        self.ev_head.record(main)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.ev_head)
            if self.peer is not None:
                self.plan.peer_adam(self.peer, n0, n, last=False)     # P2P sum + Adam of bucket 1, beside the tail GEMM
            else:
                allreduce_grads(self.flat_grad[n0:])
        if self.peer is not None and self.tail_pieces > 1:
            # the last weight-gradient GEMM in pieces of 256 gradient rows: piece i is exchanged (and its Adam done) while piece
            # i + 1 is computed, so only the last, smallest piece's exchange is exposed before the next step's forward pass
            pieces = self.plan.tail_pieces
            for i, (lo, hi) in enumerate(pieces):
                self.plan.step_finish_tail(self.features, part=i, n_parts=len(pieces))
                self.ev_piece[i].record(main)
                with torch.cuda.stream(self.comm):
                    self.comm.wait_event(self.ev_piece[i])
                    self.plan.peer_adam(self.peer, lo, hi, last=(i + 1 == len(pieces)))
            self.ev_adam.record(self.comm)
            self._adam_pending = True
            return
        self.plan.step_finish_tail(self.features)
        self.ev_bwd.record(main)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.ev_bwd)
            if self.peer is not None:
                self.plan.peer_adam(self.peer, 0, n0, last=True)
            else:
                allreduce_grads(self.flat_grad[:n0])
                self.plan.adam_step()
            self.ev_adam.record(self.comm)
        self._adam_pending = True

    def flush(self):
        if self._adam_pending:
            torch.cuda.current_stream().wait_event(self.ev_adam)
            self._adam_pending = False
