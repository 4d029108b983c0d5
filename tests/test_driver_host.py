"""Host-side pieces of the reference-compatible driver (no GPU): command line + settings overlay, adjlist conversion,
label handling, loud failure on missing dataset files."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _driver():
    spec = importlib.util.spec_from_file_location("ogl_train_main", os.path.join(ROOT, "train", "__main__.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_command_line_overlays_settings_like_the_reference():
    m = _driver()
    args, data = m.parse(["elliptic", "pytorch", "out.csv", "tsne", "--cuda", "--snapshots", "77", "--dropout", "0"])
    # file values survive unless the command line names them; store_true flags always override (reference :45-49)
    assert data["snapshots"] == 77 and data["samples"] == 45 and data["batch_timestep"] == 60 and data["priority_forward"] == 2
    assert data["cuda"] is True and data["copy_dataset_gpu"] is False and data["gpu"] == -1 and data["n_sampling_workers"] == 0
    assert data["path"] == "datasets/bitcoin" and data["embedding_size"] == 256
    with pytest.raises(SystemExit):
        m.parse(["reddit", "tf", "o.csv", "t"])            # only the pytorch backend exists here
    for ds, snaps in (("pubmed", 400), ("arxiv", 3500), ("reddit", 5000), ("elliptic", 1000)):
        assert m.parse([ds, "pytorch", "o", "t"])[1]["snapshots"] == snaps


def test_adjlist_becomes_both_directions_in_networkx_order(tmp_path):
    from ogl_b200.dataset_utils import common
    p = tmp_path / "graph.adjlist"
    p.write_text("0 2 1\n1 3\n2\n3\n4\n")                   # 0-2, 0-1, 1-3; vertex 4 isolated
    src, dst, n = common.read_adjlist_directed(str(p))
    assert n == 5
    assert list(zip(src.tolist(), dst.tolist())) == [(0, 2), (0, 1), (2, 0), (1, 0), (1, 3), (3, 1)]
    p.write_text("10 30\n30 20\n")                           # labels not 0..N-1: relabelled in sorted order
    src, dst, n = common.read_adjlist_directed(str(p))
    assert n == 3 and sorted(zip(src.tolist(), dst.tolist())) == [(0, 2), (1, 2), (2, 0), (2, 1)]


def test_labels_and_missing_files(tmp_path):
    from ogl_b200.dataset_utils import common, LOADERS
    t, labelled, n_classes = common.labels_and_classes(np.array([[0.0], [-1.0], [1.0], [1.0]]))
    assert t.dtype == np.int64 and labelled == {0, 2, 3} and n_classes == 3        # -1 counts as a class, like the reference
    assert set(LOADERS) == {"pubmed", "elliptic", "arxiv", "reddit"}
    with pytest.raises(FileNotFoundError) as e:
        common.require_files(str(tmp_path), ["feat_data.npy"])
    assert "dataset files missing" in str(e.value)
