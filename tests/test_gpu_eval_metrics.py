"""On-GPU argmax + confusion matrix + macro-F1 (ogl_eval_confusion) against sklearn on the same logits -- the host computation of
the reference's _evaluate_vertices (train/graphsage/model.py:83-86)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,C,present", [(5000, 41, 41), (300, 3, 3), (1000, 40, 7), (64, 2, 2), (1, 5, 1), (2000, 70, 70)])
def test_confusion_and_macro_f1_match_sklearn(n, C, present):
    import sklearn.metrics
    from ogl_b200 import native
    rng = np.random.default_rng(n + C)
    logits = rng.standard_normal((n, C)).astype(np.float32)
    logits[:, present:] -= 100.0                         # classes that are never predicted
    logits[rng.integers(0, n, n // 10), :] = 0.25         # rows of exact ties: numpy's argmax takes the first maximum
    labels = rng.integers(0, present, n)
    # strided logits (a slice of a wider buffer), as the eval plan produces them
    wide = torch.zeros(n, C + 3, device="cuda")
    wide[:, :C] = torch.from_numpy(logits).cuda()
    cm_dev, skipped = native.eval_confusion(wide[:, :C], torch.from_numpy(labels).cuda())
    pred = logits.argmax(axis=1)
    full = np.zeros((C, C), dtype=np.int64)
    np.add.at(full, (labels, pred), 1)
    assert int(skipped) == 0 and np.array_equal(cm_dev.cpu().numpy(), full)
    f1, cm = native.macro_f1_from_confusion(cm_dev.cpu().numpy())
    assert np.array_equal(cm, sklearn.metrics.confusion_matrix(labels, pred))
    assert abs(f1 - sklearn.metrics.f1_score(labels, pred, average="macro")) < 1e-12


def test_unknown_labels_are_skipped():
    from ogl_b200 import native
    logits = torch.randn(100, 4, device="cuda")
    labels = torch.randint(0, 4, (100,), device="cuda")
    labels[:7] = -1
    cm, skipped = native.eval_confusion(logits, labels)
    assert int(skipped) == 7 and int(cm.sum()) == 93
