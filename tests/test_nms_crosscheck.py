"""CPU: the restated tf.image.non_max_suppression (oracle/proposals.py; TF r1.6 NonMaxSuppressionV2 is not available
offline) cross-checked against an INDEPENDENT implementation of the same greedy rule that does ship in this image:
torchvision.ops.nms (suppress iff IoU > threshold, candidates by descending score).  Not a pin by TensorFlow, but a
second opinion on the selection rule, the strictness of the comparison and the zero-area behaviour.  Scores are
tie-free (std::sort in TF is unstable, torchvision's order of equal scores is its own)."""
import numpy as np
import pytest
import torch

from oracle import proposals as P

tv_ops = pytest.importorskip("torchvision.ops")


def clustered_boxes(rng, n, clusters):
    ctr = rng.uniform(0.1, 0.9, (clusters, 2))
    c = ctr[rng.integers(0, clusters, n)] + rng.normal(0, 0.02, (n, 2))
    hw = rng.uniform(0.05, 0.4, (n, 2))
    boxes = np.concatenate([c - hw / 2, c + hw / 2], -1).astype(np.float32)
    scores = rng.permutation(n).astype(np.float32) / n + np.float32(0.001)   # distinct
    return boxes, scores


@pytest.mark.parametrize("n,clusters,thr,cap", [(600, 12, 0.7, 100), (2000, 40, 0.7, 300), (300, 3, 0.45, 200),
                                                 (50, 50, 0.5, 1000), (1, 1, 0.7, 10)])
def test_oracle_nms_equals_torchvision(n, clusters, thr, cap):
    rng = np.random.default_rng(n + clusters)
    boxes, scores = clustered_boxes(rng, n, clusters)
    boxes[::17, 2:] = boxes[::17, :2]           # zero-area boxes: IoU 0 with everything, on both sides
    want = tv_ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()[:cap]
    got = P.non_max_suppression(boxes, scores, cap, thr)
    fast = P.non_max_suppression_fast(boxes, scores, cap, thr)
    assert np.array_equal(np.asarray(got), want) and np.array_equal(np.asarray(fast), want)
    assert len(want) > 0 and (n < 100 or len(want) < n)


def test_threshold_is_strict():
    """IoU exactly equal to the threshold keeps the box (TF: suppressed iff iou > threshold), on both sides."""
    boxes = np.array([[0, 0, 1, 1], [0, 0.5, 1, 1.5]], np.float32)      # IoU = 0.5 / 1.5 = 1/3
    scores = np.array([0.9, 0.8], np.float32)
    iou = np.float32(0.5) / np.float32(1.5)
    for thr, kept in ((float(iou), 2), (float(np.nextafter(iou, np.float32(0))), 1)):
        want = tv_ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()
        assert len(want) == kept
        assert np.array_equal(np.asarray(P.non_max_suppression(boxes, scores, 10, thr)), want)


def test_top_k_tie_rule_against_torch_stable_sort():
    """tf.nn.top_k (TopKV2, not available offline) is documented as: descending, equal values -> the lower index
    first.  The oracle states it as a lexsort; torch's stable descending sort is an independent implementation of the
    same rule (scores with many exact ties, +-0.0 and a few +-inf)."""
    rng = np.random.default_rng(11)
    for n, k in ((5000, 600), (64, 64), (7, 3), (1, 1)):
        s = (rng.integers(0, 50, n) / 50.0).astype(np.float32)          # ~100 copies of each value
        s[rng.integers(0, n, max(n // 50, 1))] = np.float32(-0.0)
        if n > 10:
            s[3], s[n // 2] = np.inf, -np.inf
        boxes = rng.random((n, 4), dtype=np.float32)
        order = np.lexsort((np.arange(n), -s.astype(np.float64)))[:k]
        want = torch.sort(torch.from_numpy(s), descending=True, stable=True).indices.numpy()[:k]
        assert np.array_equal(order, want)
        # the same rule inside _filter_and_sort_boxes (everything passes the filter: size 1, centre inside)
        b = np.tile(np.array([0.25, 0.25, 0.75, 0.75], np.float32), (n, 1)) + boxes * np.float32(1e-3)
        pos = np.where(np.isfinite(s) & (s > 0), s, np.float32(0.5)).astype(np.float32)
        ss, bb, sel = P.filter_and_sort_boxes(pos, b, 0.01, k)
        want = torch.sort(torch.from_numpy(pos), descending=True, stable=True).indices.numpy()[:k]
        assert np.array_equal(sel, want) and np.array_equal(ss[:len(want)], pos[want])
        assert np.array_equal(bb[:len(want)], b[want])
