"""Write a synthetic dataset in the reference's on-disk formats so that `python train <dataset> pytorch ...`
can run without the (undownloadable) real data:

  vertex streams (pubmed / elliptic / arxiv):  feat_data.npy | feats.npy (float64 [V, F]), targets.npy (float64 [V, 1], -1 =
      unlabelled), graph.adjlist (networkx), vertex_timestamp.json | postponed_timestamp.json ({vertex: time})
  edge stream (reddit):  feat_data.npy, targets.npy, edges_dataframe.csv (src,dst; time ordered; ids dense in first-appearance order)

    python tools/make_synthetic_dataset.py pubmed /tmp/pubmed --vertices 2000 --edges 9000 --feats 50 --classes 3
"""
import argparse
import json
import os

import numpy as np


def planted(V, E, F, C, rng):
    y = rng.integers(0, C, V)
    src = rng.integers(0, V, E)
    cand = rng.integers(0, V, (E, 8))
    pick = np.argmax(y[cand] == y[src][:, None], axis=1)
    dst = np.where(rng.random(E) < 0.85, cand[np.arange(E), pick], rng.integers(0, V, E))
    x = rng.standard_normal((V, F)) * 0.7
    x[np.arange(V), y % F] += 1.5
    return src, dst, x, y.astype(np.float64).reshape(-1, 1)


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("dataset", choices=["pubmed", "elliptic", "arxiv", "reddit"])
    ap.add_argument("out")
    ap.add_argument("--vertices", type=int, default=2000)
    ap.add_argument("--edges", type=int, default=9000)
    ap.add_argument("--feats", type=int, default=50)
    ap.add_argument("--classes", type=int, default=3)
    ap.add_argument("--unlabelled", type=float, default=0.0)
    ap.add_argument("--seed", type=int, default=0)
    a = ap.parse_args(argv)
    rng = np.random.default_rng(a.seed)
    os.makedirs(a.out, exist_ok=True)
    src, dst, x, y = planted(a.vertices, a.edges, a.feats, a.classes, rng)
    if a.dataset == "reddit":
        order = {}
        for u, v in zip(src.tolist(), dst.tolist()):
            for w in (u, v):
                if w not in order:
                    order[w] = len(order)
        perm = np.array(sorted(order, key=order.get))
        m = np.full(a.vertices, -1)
        m[perm] = np.arange(len(perm))
        x, y = x[perm], y[perm]
        if a.unlabelled:
            y[rng.random(len(y)) < a.unlabelled] = -1
        np.save(os.path.join(a.out, "feat_data.npy"), x)
        np.save(os.path.join(a.out, "targets.npy"), y)
        with open(os.path.join(a.out, "edges_dataframe.csv"), "w") as f:
            f.write("src,dst\n")
            for u, v in zip(m[src].tolist(), m[dst].tolist()):
                f.write("%d,%d\n" % (u, v))
        return
    import networkx as nx
    G = nx.Graph()
    G.add_nodes_from(range(a.vertices))           # every vertex present, ids 0..V-1
    G.add_edges_from(zip(src.tolist(), dst.tolist()))
    nx.write_adjlist(G, os.path.join(a.out, "graph.adjlist"))
    if a.unlabelled:
        y[rng.random(len(y)) < a.unlabelled] = -1
    np.save(os.path.join(a.out, "feats.npy" if a.dataset == "arxiv" else "feat_data.npy"), x)
    np.save(os.path.join(a.out, "targets.npy"), y)
    ts = {int(v): float(t) for v, t in enumerate(rng.permutation(a.vertices))}
    with open(os.path.join(a.out, "postponed_timestamp.json" if a.dataset == "pubmed" else "vertex_timestamp.json"), "w") as f:
        json.dump(ts, f)


if __name__ == "__main__":
    main()
