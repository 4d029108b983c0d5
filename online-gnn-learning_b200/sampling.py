"""dgl.sampling surface used by the reference trainers (train/graphsage/pytorch/model.py:44-47),
served by the on-GPU Philox sampler + to_block kernels of one ogl_plan."""
import torch

NID = "_ID"
EID = "_ID"


class Block:
    """One bipartite sampled block (fixed-fanout ELL layout on the device)."""
    is_block = True

    def __init__(self, plan, hop):
        self._plan, self._hop, self._stamp = plan, hop, plan._stamp
        self._dst = plan.level_nodes(hop)
        self._src = plan.level_nodes(hop + 1)
        self._lid, self._gsrc, self._eid, self.fanout = plan.block_edges(hop)
        self.srcdata = {NID: self._src.long()}
        self.dstdata = {NID: self._dst.long()}
        valid = self._lid >= 0
        self.edata = {EID: self._eid[valid]}
        self._valid = valid

    def to(self, device):
        return self

    def number_of_dst_nodes(self):
        return self._dst.numel()

    def number_of_src_nodes(self):
        return self._src.numel()

    def number_of_edges(self):
        return int(self._valid.sum().item())

    def edges(self):
        """(src_local, dst_local) in DGL order: by dst position, then pick index."""
        dst = torch.arange(self._dst.numel(), device="cuda").repeat_interleave(self.fanout)
        return self._lid[self._valid].long(), dst[self._valid]


class MultiLayerNeighborSampler:
    """fanouts listed input-layer first like DGL; the reference passes [samples]*2, replace=True."""

    def __init__(self, fanouts, replace=True, return_eids=True):
        if not replace:
            raise NotImplementedError("the reference samples with replacement (pytorch/model.py:44)")
        self.fanouts = list(fanouts)
        self.hop_fanouts = list(reversed(self.fanouts))     # hop 0 = seeds hop


class NodeDataLoader:
    """Iterates (input_nodes, seeds, blocks) minibatches in order (shuffle=True permutes seeds with torch's RNG
    like DataLoader does)."""

    def __init__(self, graph, nids, sampler, batch_size, shuffle=False, drop_last=False, num_workers=0, plan=None):
        self.graph, self.sampler, self.batch_size = graph, sampler, max(1, int(batch_size))
        self.nids = torch.as_tensor(nids, dtype=torch.int64)
        self.shuffle, self.drop_last = shuffle, drop_last
        self.plan = plan if plan is not None else graph.sampling_plan(sampler.hop_fanouts, self.batch_size)

    def __len__(self):
        n = self.nids.numel()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        nids = self.nids
        if self.shuffle:
            nids = nids[torch.randperm(nids.numel())]
        nids = nids.cuda()
        for i in range(0, nids.numel(), self.batch_size):
            seeds = nids[i:i + self.batch_size]
            if self.drop_last and seeds.numel() < self.batch_size:
                break
            self.plan.sample(self.graph.native, seeds)
            L = self.plan.L
            blocks = [Block(self.plan, hop) for hop in reversed(range(L))]
            yield blocks[0].srcdata[NID], seeds, blocks
