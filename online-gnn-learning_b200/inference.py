"""Streaming inference with cached per-vertex intermediates on the GPU: the drop-in for the per-request method of the reference's
TorchServe handler (`MNISTDigitClassifier.inference`, /root/reference/inference_optimized.py:144-301).

The serving graph lives in one streaming CSR of the C library that stores every edge in both directions (rows [0, V): in-edge
sources, rows [V, 2V): out-edge targets, whose degrees are the out-degrees); the
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
        # ONE streaming CSR holds both directions: row v = sources of the stored edges u -> v, row v_cap + u = their targets
        # (degree of row v_cap + u = out-degree of u); a request's edges go in with one insert call
        self.g = native.Graph(2 * self.v_cap, 2 * e_cap)
        self.g.insert_vertices(2 * self.v_cap)
        self.cap_in, self.cap_out = 4096, 1024
        self._q = torch.zeros(5 + 3 * 256 + self.cap_in + 2 * self.cap_out, dtype=torch.int64, device="cuda")
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
        self.n = max(self.n, n_new)                               # new vertices: zero caches (already), dataset feature rows (resident)
        pr = np.asarray(pairs, dtype=np.int64)
        a, b = pr[:, 0], pr[:, 1]
        # stored reversed, b -> a (:181-182): a's in-row gains b, b's out-row (row v_cap + b) gains a
        self.g.insert_edges(np.concatenate([b, a]), np.concatenate([a, b + self.v_cap]), symmetric=False)
        l_all = np.array(list(vertices), dtype=np.int64)
        v0, pred, succs = self._query(l_all)
        P = list(set(pred))
        S = list(set(succs))
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
                native.induced_mean(self.g, self.member, sub_t, c[proj], c[neigh])
                self.member[sub_t] = 0
            native.rows_linear(x, ids, self.w[L + "fc_self.weight"], self.w[L + "fc_self.bias"], c[out], ids, relu=(i < 1),
                               x2=c[neigh], ids2=ids, w2=self.w[L + "fc_neigh.weight"], b2=self.w[L + "fc_neigh.bias"])
        self.requests += 1
        if not P:
            return P, []
        classes = c["h2"][torch.as_tensor(P, dtype=torch.int64).cuda()].argmax(dim=1).cpu().tolist()
        return P, classes

    def _query(self, l_all):
        """(V0, pred, succs) of the request's vertices (inference_optimized.py:185-211): one launch + one D2H copy; the multi-call
        path takes over when a hub's rows exceed the scratch"""
        n = len(l_all)
        lv = torch.as_tensor(l_all).cuda()
        if n <= 256:
            native.infer_query(self.g, lv, self.v_cap, TH, self.cap_in, self.cap_out, self._q)
            q = self._q.cpu().numpy()
            if q[0] == 0:
                deg = q[3:3 + n]
                keep = l_all[deg < TH]
                base_in = 3 + n + 2 * (n + 1)
                pred = q[base_in:base_in + q[1]].tolist()
                base_out = base_in + self.cap_in
                dst = q[base_out:base_out + q[2]]
                dst_deg = q[base_out + self.cap_out:base_out + self.cap_out + q[2]]
                return list(set(keep.tolist())), pred, dst[dst_deg < TH].tolist()
        out_deg = self.g.row_degrees(lv + self.v_cap).cpu().numpy()
        keep = l_all[out_deg < TH]
        kv = torch.as_tensor(keep).cuda()
        _, succs = self.g.gather_rows(kv + self.v_cap)
        _, pred = self.g.gather_rows(kv)
        if succs.numel():
            succs = succs[self.g.row_degrees(succs + self.v_cap) < TH]
        return list(set(keep.tolist())), pred.cpu().tolist(), succs.cpu().tolist()

    # ---- the handler's call surface (inference_optimized.py:144, :304-318) --------------------------------------------
    def inference(self, data):
        value = json.loads(str(data[0].get("body")))
        _, classes = self.request(value)
        return [str(classes)]
