"""CPU oracle for the streaming GraphSAGE hot path -- TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (numpy / torch-CPU, pure-Python loops only on
small cases) of the reference algorithm for the path named by
BASELINE.json:north_star.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product package (``online-gnn-learning_b200/``) never imports it and has no CPU
fallback.

Pinning status (see DESIGN.md section "Oracle"):
  * ``sumtree`` / ``replay``   -- PINNED: checked bit-for-bit against the
    reference's own ``train/prioritized_replay/*.py`` and
    ``train/graph/train_test_graph.py`` executed in the build container
    (fixtures under tests/golden/, generator tests/golden/make_golden.py).
  * ``graph`` (edge / vertex streams) -- PINNED against the reference's own
    ``train/graph/dynamic_graph_{edge,vertex}.py`` run over a recording mini-dgl
    shim (same generator script).
  * ``sampler`` / ``to_block`` / ``sage`` (the DGL-executed pieces) -- PARITY
    UNPINNED: DGL 0.5.x is un-vendored, un-pinned and absent, and the reference
    holds no tests or golden vectors at that boundary.  The restatement follows
    the reference call sites and the formula evidenced by
    ``inference_optimized.py:135-139,258-260,273-276``; Philox4x32-10 itself is
    pinned against the Random123 known-answer vectors.
"""
