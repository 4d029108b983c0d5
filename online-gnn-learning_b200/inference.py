"""Streaming inference with cached per-vertex intermediates on the GPU: the drop-in for the per-request method of the reference's
TorchServe handler (`MNISTDigitClassifier.inference`, /root/reference/inference_optimized.py:144-301).

The serving graph lives in two streaming CSRs of the C library (in-edges, and the reversed copy for out-edges / out-degrees); the
per-vertex caches h0proj / neigh0 / h1 / h1proj / neigh1 / h2, the feature table and the weights stay resident in HBM; the dense
row updates and the induced-subgraph mean are csrc/infer.cu kernels (fp32, as the reference serves).  The request's vertex sets are
tiny (bounded by the handler's out-degree threshold of 15) and the handler's ANSWER ORDER is CPython's set order, so the sets are
formed on the host from the few adjacency entries the device returns -- exactly as the reference forms them.

Every quirk of the handler is kept (reversed edge storage, mean instead of the training-time max, neigh rewritten for every vertex
of the predecessor set, out-degree filters, answer = cached h2 of the predecessor set); tests/test_gpu_inference.py checks this class
against oracle/inference.py, which is pinned to the reference's own code."""
import json

import numpy as np
import torch

from . import _native as native

TH = 15   # sampling_th (inference_optimized.py:184)


class CachedInference:
    def __init__(self, feat, state_dict, v_cap=None, e_cap=None):
        """feat: [V, F] dataset features (array / tensor); state_dict: `layers.{0,1}.{fc_pool,fc_self,fc_neigh}.{weight,bias}` as
        saved by the reference's export_model.py / GraphSAGE.state_dict()"""
        self.feat = torch.as_tensor(np.asarray(feat, dtype=np.float32) if not isinstance(feat, torch.Tensor) else feat).to("cuda", torch.float32).contiguous()
        V, F = self.feat.shape
        self.v_cap = int(v_cap or V)
        e_cap = int(e_cap or max(16 * self.v_cap, 1 << 16))
        self.w = {k: torch.as_tensor(np.asarray(v) if not isinstance(v, torch.Tensor) else v).detach().to("cuda", torch.float32).contiguous()
                  for k, v in state_dict.items() if k.startswith("layers.")}
        H = self.w["layers.0.fc_self.weight"].shape[0]
        C = self.w["layers.1.fc_self.weight"].shape[0]
        self.g_in = native.Graph(self.v_cap, e_cap)     # row v: sources of the stored edges u -> v
        self.g_out = native.Graph(self.v_cap, e_cap)    # row u: targets of the stored edges u -> v  (degree = out-degree)
        dims = dict(h0proj=F, neigh0=F, h1=H, h1proj=H, neigh1=H, h2=C)
        self.cache = {k: torch.zeros(self.v_cap, d, device="cuda") for k, d in dims.items()}
        self.member = torch.zeros(self.v_cap, dtype=torch.uint8, device="cuda")
        self.n = 0
        self.requests = 0

    def __len__(self):
        return self.n

    # ---- one request ---------------------------------------------------------------------------------------------
    def request(self, pairs):
        """pairs: [[a, b], ...] -> (P, classes): the vertices answered for (the handler's order) and their predicted classes"""
        vertices, total = set(), set()
        for a, b in pairs:
            vertices.add(a)
            total.add(a)
            total.add(b)
        n_new = max(total) + 1
        assert n_new <= self.v_cap, "vertex id %d beyond the feature table / capacity %d" % (n_new - 1, self.v_cap)
        if n_new > self.n:                                        # new vertices: zero caches (already), dataset feature rows (resident)
            self.g_in.insert_vertices(n_new - self.n)
            self.g_out.insert_vertices(n_new - self.n)
            self.n = n_new
        pr = torch.as_tensor(np.asarray(pairs, dtype=np.int64)).cuda()
        a, b = pr[:, 0].contiguous(), pr[:, 1].contiguous()
        self.g_in.insert_edges(b, a, symmetric=False)             # stored reversed: b -> a (:181-182)
        self.g_out.insert_edges(a, b, symmetric=False)
        l_vertices = np.array(list(vertices), dtype=np.int64)
        out_deg = self.g_out.row_degrees(l_vertices).cpu().numpy()
        l_vertices = l_vertices[out_deg < TH]
        v0 = list(set(l_vertices.tolist()))
        lv = torch.as_tensor(l_vertices).cuda()
        _, succs = self.g_out.gather_rows(lv)
        _, pred = self.g_in.gather_rows(lv)
        P = list(set(pred.cpu().tolist()))
        if succs.numel():
            keep = self.g_out.row_degrees(succs) < TH
            succs = succs[keep]
        S = list(set(succs.cpu().tolist()))
        self.last_sets = (v0, P, S)
        c = self.cache
        for i, (nids, sub, x, proj, neigh, out) in enumerate(((v0, P, self.feat, "h0proj", "neigh0", "h1"), (S, S, c["h1"], "h1proj", "neigh1", "h2"))):
            if not nids:
                continue
            ids = torch.as_tensor(nids, dtype=torch.int64).cuda()
            L = "layers.%d." % i
            native.rows_linear(x, ids, self.w[L + "fc_pool.weight"], self.w[L + "fc_pool.bias"], c[proj], ids, relu=True)
            if sub:
                sub_t = ids if sub is nids else torch.as_tensor(sub, dtype=torch.int64).cuda()
                self.member[sub_t] = 1
                native.induced_mean(self.g_in, self.member, sub_t, c[proj], c[neigh])
                self.member[sub_t] = 0
            native.rows_linear(x, ids, self.w[L + "fc_self.weight"], self.w[L + "fc_self.bias"], c[out], ids, relu=(i < 1),
                               x2=c[neigh], ids2=ids, w2=self.w[L + "fc_neigh.weight"], b2=self.w[L + "fc_neigh.bias"])
        self.requests += 1
        if not P:
            return P, []
        classes = c["h2"][torch.as_tensor(P, dtype=torch.int64).cuda()].argmax(dim=1).cpu().tolist()
        return P, classes

    # ---- the handler's call surface (inference_optimized.py:144, :304-318) --------------------------------------------
    def inference(self, data):
        value = json.loads(str(data[0].get("body")))
        _, classes = self.request(value)
        return [str(classes)]
