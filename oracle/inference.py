"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's cached streaming inference, one request at a time:
`MNISTDigitClassifier.inference`, /root/reference/inference_optimized.py:144-301.  Plain numpy fp32 + Python sets / lists.

PINNED: tests/test_oracle_pinning.py checks it against tests/golden/inference_stream.npz, which was produced by running the
reference's own method (tests/golden/make_golden_inference.py).

What the method does per request (a JSON list of [a, b] pairs), all quirks kept:
  * the graph grows to max id + 1 vertices; new vertices get zero caches and their dataset feature rows (:157-176)
  * every pair is stored REVERSED, as the edge b -> a (:181-182)
  * V0 = the a's whose out-degree is < 15 (:186-192); P = sources of V0's in-edges (:196, :204); S = targets of V0's out-edges whose
    out-degree is < 15 (:194-195, :206-211)
  * layer 0 on V0: h0proj = relu(fc_pool0(feat)); neigh0 is rewritten for EVERY vertex of P as the MEAN (not the training-time max)
    of h0proj over the in-edges that lie inside P, 0 without one (:255-276, :268); h1 = relu(fc_self0(feat) + fc_neigh0(neigh0))
  * layer 1 on S: the same with h1 / h1proj / neigh1 over the in-edges inside S, no ReLU -> h2 (:247-286)
  * answer = argmax of the cached h2 of the vertices of P, in P's order (:289)
Set / list orders follow CPython's (the reference builds them the same way), so the answer order is reproduced too."""
import numpy as np

TH = 15   # sampling_th, inference_optimized.py:184


class CachedInferenceOracle:
    def __init__(self, feat, params):
        """feat: [V, F] fp32 dataset features; params: {"layers.{l}.{fc_pool,fc_self,fc_neigh}.{weight,bias}": ndarray} for l = 0, 1"""
        self.feat_all = np.asarray(feat, dtype=np.float32)
        self.w = {k: np.asarray(v, dtype=np.float32) for k, v in params.items()}
        F = self.feat_all.shape[1]
        H = self.w["layers.0.fc_self.weight"].shape[0]
        C = self.w["layers.1.fc_self.weight"].shape[0]
        self.dims = dict(h0proj=F, neigh0=F, h1=H, h1proj=H, neigh1=H, h2=C)
        self.n = 0
        self.src, self.dst = [], []
        self.cache = {k: np.zeros((0, d), dtype=np.float32) for k, d in self.dims.items()}
        self.feat = np.zeros((0, F), dtype=np.float32)

    def _lin(self, l, name, x):
        return x @ self.w["layers.%d.%s.weight" % (l, name)].T + self.w["layers.%d.%s.bias" % (l, name)]

    def _grow(self, n_new):
        if n_new <= self.n:
            return
        add = n_new - self.n
        for k, d in self.dims.items():
            self.cache[k] = np.concatenate([self.cache[k], np.zeros((add, d), dtype=np.float32)])
        self.feat = np.concatenate([self.feat, self.feat_all[self.n:n_new]])
        self.n = n_new

    def _induced_mean(self, nodes, proj):
        """mean of proj[src] over the edges src -> v with src and v in `nodes` (parallel edges counted), rows in `nodes` order"""
        member = set(nodes)
        out = np.zeros((len(nodes), proj.shape[1]), dtype=np.float32)
        where = {v: i for i, v in enumerate(nodes)}
        msgs = {}
        for a, b in zip(self.src, self.dst):
            if a in member and b in member:
                msgs.setdefault(b, []).append(a)
        for v, srcs in msgs.items():
            out[where[v]] = proj[np.asarray(srcs)].mean(axis=0, dtype=np.float32)
        return out

    def request(self, pairs):
        """-> (P, classes): the vertices answered for and their predicted classes; also updates the caches"""
        vertices, total = set(), set()
        for a, b in pairs:
            vertices.add(a)
            total.add(a)
            total.add(b)
        self._grow(max(total) + 1)
        for a, b in pairs:
            self.src.append(int(b))
            self.dst.append(int(a))
        src, dst = np.asarray(self.src), np.asarray(self.dst)
        out_deg = np.bincount(src, minlength=self.n)
        l_vertices = np.array(list(vertices))
        l_vertices = l_vertices[out_deg[l_vertices] < TH]
        v0 = set(l_vertices.tolist())
        succs, pred = [], []
        for x in l_vertices.tolist():
            succs += dst[src == x].tolist()
        for x in l_vertices.tolist():
            pred += src[dst == x].tolist()
        P = list(set(pred))
        succs = np.asarray(succs, dtype=np.int64)
        succs = succs[out_deg[succs] < TH] if len(succs) else succs
        S = list(set(succs.tolist()))
        self.last_sets = (list(v0), P, S)
        for i in range(2):
            nids = list(v0) if i == 0 else list(S)
            h_self = self.feat[nids] if i == 0 else self.cache["h1"][nids]
            proj, neigh, out = ("h0proj", "neigh0", "h1") if i == 0 else ("h1proj", "neigh1", "h2")
            self.cache[proj][nids] = np.maximum(self._lin(i, "fc_pool", h_self), 0.0)
            sub = P if i == 0 else S
            if len(sub):
                self.cache[neigh][sub] = self._induced_mean(sub, self.cache[proj])
            rst = self._lin(i, "fc_self", h_self) + self._lin(i, "fc_neigh", self.cache[neigh][nids])
            if i < 1:
                rst = np.maximum(rst, 0.0)
            self.cache[out][nids] = rst
        classes = self.cache["h2"][P].argmax(axis=1).tolist() if len(P) else []
        return P, classes
