"""`python train <dataset> pytorch out.csv tsne --cuda` (the reference's command line) end to end over
synthetic datasets written in the reference's on-disk formats (8(f)-2): vertex stream (adjlist + timestamps) and edge
stream (csv)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, **kw):
    r = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=600, **kw)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return r


@pytest.mark.parametrize("dataset,extra", [("pubmed", ["--unlabelled", "0.2"]), ("reddit", []), ("elliptic", ["--unlabelled", "0.5", "--classes", "2"])])
def test_reference_command_line(tmp_path, dataset, extra):
    data = str(tmp_path / dataset)
    _run(["tools/make_synthetic_dataset.py", dataset, data, "--vertices", "1500", "--edges", "8000", "--feats", "24"] + extra)
    out = str(tmp_path / "res.csv")
    _run(["train", dataset, "pytorch", out, str(tmp_path / "tsne"), "--cuda", "--path", data, "--snapshots", "20", "--delta", "2",
          "--eval", "3", "--batch_timestep", "4", "--batch_size", "32", "--embedding_size", "16", "--samples", "5", "--train_offline", "4",
          "--epochs_offline", "1", "--batch_full", "256", "--max_timesteps", "12"])
    rows = [l for l in open(out).read().strip().split("\n") if l]
    models = [r.split(";")[0] for r in rows]
    assert set(models) == {"random", "prioritized", "no_rehersal", "offline"}
    assert all(len(r.split(";")) == 4 for r in rows)
    scored = [float(r.split(";")[1]) for r in rows if r.split(";")[1]]
    assert len(scored) >= 8 and all(0.0 <= s <= 1.0 for s in scored)


def test_missing_dataset_fails_loudly(tmp_path):
    r = subprocess.run([sys.executable, "train", "pubmed", "pytorch", str(tmp_path / "o.csv"), "x", "--cuda", "--path", str(tmp_path / "nope")],
                       cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "dataset files missing" in r.stderr
