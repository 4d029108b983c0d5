"""Cached streaming inference on the GPU (online-gnn-learning_b200/inference.py + csrc/infer.cu) against the oracle restatement of the
reference's handler method (oracle/inference.py, pinned to the reference's own code by tests/test_oracle_pinning.py)."""
import json

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from test_oracle_pinning import inference_fixture  # noqa: E402


def check_stream(feat, params, requests, tag):
    from ogl_b200.inference import CachedInference
    from oracle.inference import CachedInferenceOracle
    o = CachedInferenceOracle(feat, params)
    d = CachedInference(feat, {k: torch.from_numpy(np.asarray(v)) for k, v in params.items()})
    for r, pairs in enumerate(requests):
        P, classes = o.request(pairs)
        got = json.loads(d.inference([{"body": json.dumps(pairs)}])[0])
        assert d.last_sets[1] == P and d.last_sets[2] == o.last_sets[2] and sorted(d.last_sets[0]) == sorted(o.last_sets[0]), \
            "%s request %d: vertex sets" % (tag, r)
        assert len(d) == o.n
        for k in o.cache:
            want = o.cache[k]
            have = d.cache[k][:o.n].cpu().numpy()
            scale = max(1.0, float(np.abs(want).max()))
            assert np.abs(have - want).max() <= 1e-5 * scale, "%s request %d: cache %s differs by %g" % (tag, r, k, np.abs(have - want).max())
        # classes: equal wherever the oracle's top-2 logits are not within rounding of each other
        h2 = o.cache["h2"][P]
        for v, cw, cg, row in zip(P, classes, got, h2):
            top = np.sort(row)[::-1]
            if top[0] - top[1] > 1e-4 * max(1.0, abs(top[0])):
                assert cw == cg, "%s request %d vertex %d: class %d vs %d" % (tag, r, v, cg, cw)


def test_cached_inference_on_the_reference_fixture(golden):
    """the very request stream the reference's own handler was run on (tests/golden/inference_stream.npz)"""
    feat, params, reqs = inference_fixture(golden("inference_stream"))
    check_stream(feat, params, [q["pairs"] for q in reqs], "fixture")
    # and the recorded answers of the reference itself
    from ogl_b200.inference import CachedInference
    d = CachedInference(feat, {k: torch.from_numpy(v) for k, v in params.items()})
    for r, q in enumerate(reqs):
        P, classes = d.request(q["pairs"])
        assert P == q["P"], "request %d: answered vertices" % r
        agree = sum(int(a == b) for a, b in zip(classes, q["out"]))
        assert agree >= len(P) - 1, "request %d: %d of %d classes agree with the reference" % (r, agree, len(P))


def test_cached_inference_larger_stream():
    """600 requests over 3000 vertices, 166 features (Elliptic-shaped layer widths), hubs crossing the out-degree threshold"""
    rng = np.random.default_rng(21)
    V, F, H, C = 3000, 166, 64, 2
    feat = rng.standard_normal((V, F)).astype(np.float32)
    params = {}
    for l, (i, o) in enumerate(((F, H), (H, C))):
        for name, (oo, ii) in (("fc_pool", (i, i)), ("fc_self", (o, i)), ("fc_neigh", (o, i))):
            params["layers.%d.%s.weight" % (l, name)] = (rng.standard_normal((oo, ii)) / np.sqrt(ii)).astype(np.float32)
            params["layers.%d.%s.bias" % (l, name)] = (0.1 * rng.standard_normal(oo)).astype(np.float32)
    reqs, hi = [], 20
    for r in range(600):
        hi = min(V, hi + int(rng.integers(0, 12)))
        k = int(rng.integers(1, 6))
        pairs = []
        for _ in range(k):
            a, b = int(rng.integers(0, hi)), int(rng.integers(0, hi))
            pairs.append([a, b])
            if rng.random() < 0.5:
                pairs.append([b, a])
        if r % 4 == 0:
            pairs.append([int(rng.integers(0, hi)), int(rng.integers(0, 5))])      # hubs 0..4 gain out-edges
        reqs.append(pairs)
    check_stream(feat, params, reqs, "large")
