"""CPU: VOC metric arithmetic (x-detector_b200/utility/metrics.py) against hand-computed cases and the independent
devkit-style AP of oracle/voc_eval.py."""
import numpy as np

import xdet_b200  # noqa: F401
from oracle import voc_eval as ov
from xdet_b200.utility import metrics as M


def test_ap_hand_case():
    # 4 detections sorted by score: TP, FP, TP, FP with 2 objects -> precision 1, .5, .667, .5; recall .5, .5, 1, 1
    p, r = M.precision_recall(2, 4, [1, 0, 1, 0], [0, 1, 0, 1], [0.9, 0.8, 0.7, 0.6])
    assert np.allclose(p, [1, 0.5, 2 / 3, 0.5]) and np.allclose(r, [0.5, 0.5, 1, 1])
    assert abs(M.average_precision_voc12(p, r) - (0.5 * 1 + 0.5 * 2 / 3)) < 1e-12
    assert abs(M.average_precision_voc07(p, r) - (6 * 1 + 5 * 2 / 3) / 11) < 1e-12


def test_ap_matches_devkit_formula():
    rng = np.random.default_rng(0)
    for _ in range(20):
        n = int(rng.integers(1, 200))
        tp = rng.random(n) < 0.4
        fp = ~tp
        scores = rng.random(n).astype(np.float32)
        nobj = int(tp.sum() + rng.integers(0, 20))
        p, r = M.precision_recall(nobj, n, tp, fp, scores)
        assert abs(M.average_precision_voc07(p, r) - ov.voc_ap(r, p, True)) < 1e-12
        assert abs(M.average_precision_voc12(p, r) - ov.voc_ap(r, p, False)) < 1e-12


def test_streaming_accumulation_and_filters():
    st = M.streaming_tp_fp_arrays({1: [2]}, {1: [[1, 0, 0]]}, {1: [[0, 1, 1]]}, {1: [[0.9, 0.5, 0.0]]})
    st = M.streaming_tp_fp_arrays({1: [1]}, {1: [[0, 0]]}, {1: [[0, 0]]}, {1: [[0.7, 0.6]]}, state=st)
    nobj, ndet, tp, fp, sc = st[1].value()
    # the zero-score FP and the two neither-TP-nor-FP ('difficult') detections are dropped
    assert (nobj, ndet) == (3, 2) and tp.tolist() == [True, False] and fp.tolist() == [False, True]
    m, aps = M.voc_map(st)
    assert 0.0 <= m <= 1.0 and list(aps) == [1]


def test_ap_matches_the_reference_voc_ap_golden():
    """Golden vectors minted from the reference's own numpy voc_ap (voc_eval.py:98-130, executed unmodified by
    tests/golden/make_voc_ap_golden.py): utility/metrics.py's precision/recall + AP07/AP12 reproduce them."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voc_ap_golden.npz"))
    for i in range(int(g["n_cases"])):
        tp = g["tp_%d" % i]
        n = tp.shape[0]
        scores = np.linspace(1.0, 0.5, n).astype(np.float32)       # already in score order
        p, r = M.precision_recall(int(g["npos_%d" % i]), n, tp, ~tp, scores)
        assert abs(M.average_precision_voc07(p, r) - float(g["ap07_%d" % i])) < 1e-12, i
        assert abs(M.average_precision_voc12(p, r) - float(g["ap12_%d" % i])) < 1e-12, i
        assert abs(ov.voc_ap(r, p, True) - float(g["ap07_%d" % i])) < 1e-12
        assert abs(ov.voc_ap(r, p, False) - float(g["ap12_%d" % i])) < 1e-12


def test_streaming_metrics_match_reference_metrics_py():
    """Goldens from the reference's own utility/metrics.py run under the numpy TensorFlow stand-in
    (tests/golden/make_metrics_golden.py), chained as bboxes_eval chains them: accumulate over 7 images and 4 classes
    (zero-padded scores, sub-threshold scores, neither-TP-nor-FP detections, a never-correct class), then
    precision / recall, AP07, AP12 per class and the mAPs."""
    import os
    G = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics_golden.npz"))
    classes = [int(c) for c in G["classes"]]
    state = None
    for i in range(int(G["images"])):
        args = [{c: G["in_%d_%d_%s" % (i, c, k)] for c in classes} for k in ("n", "tp", "fp", "scores")]
        state = M.streaming_tp_fp_arrays(*args, state=state)
    for c in classes:
        nobj, ndet, tp, fp, sc = state[c].value()
        assert (nobj, ndet) == (int(G["acc_%d_nobjects" % c]), int(G["acc_%d_ndetections" % c]))
        assert np.array_equal(tp, G["acc_%d_tp" % c]) and np.array_equal(fp, G["acc_%d_fp" % c])
        assert np.array_equal(sc.view(np.int32), G["acc_%d_scores" % c].view(np.int32))
        p, r = M.precision_recall(nobj, ndet, tp, fp, sc)
        assert np.abs(p - G["prec_%d" % c]).max() < 1e-15 and np.abs(r - G["rec_%d" % c]).max() < 1e-15
        assert abs(M.average_precision_voc07(p, r) - float(G["ap07_%d" % c])) < 1e-12
        assert abs(M.average_precision_voc12(p, r) - float(G["ap12_%d" % c])) < 1e-12
    assert abs(M.voc_map(state, True)[0] - float(G["map07"])) < 1e-12
    assert abs(M.voc_map(state, False)[0] - float(G["map12"])) < 1e-12
