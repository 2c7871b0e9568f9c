"""CPU oracle (TEST INFRASTRUCTURE ONLY) for the TP/FP matching and the VOC AP arithmetic of the reference's eval
path: loop-level restatement of ``bboxes_jaccard`` / ``bboxes_matching`` (utility/eval_helper.py:671-788) in numpy
fp32, and an independent textbook VOC AP (the well-known ``voc_ap`` of the PASCAL devkit / py-faster-rcnn) used to
cross-check ``utility/metrics.py``'s formulas.  Pinned: bboxes_matching by the reference's own eval_helper.py under
the numpy TensorFlow stand-in (tests/golden/tfpath_golden.npz), the AP arithmetic by the reference's own
voc_eval.py:98-130 (tests/golden/voc_ap_golden.npz)."""
import numpy as np

F = np.float32


def bboxes_jaccard(bbox_ref, bboxes):
    """utility/eval_helper.py:671-699 (fp32, same operation order)."""
    b = bboxes.astype(F)
    r = np.asarray(bbox_ref, F)
    h = np.maximum(np.minimum(b[:, 2], r[2]) - np.maximum(b[:, 0], r[0]), F(0))
    w = np.maximum(np.minimum(b[:, 3], r[3]) - np.maximum(b[:, 1], r[1]), F(0))
    inter = (h * w).astype(F)
    union = ((-inter + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])).astype(F) + F((r[2] - r[0]) * (r[3] - r[1]))).astype(F)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(union > 0, inter / union, F(0)).astype(F)


def bboxes_matching(label, scores, bboxes, glabels, gbboxes, gdifficults, matching_threshold=0.5):
    """utility/eval_helper.py:700-788 for one image and one class."""
    glabels = np.asarray(glabels)
    gdiff = np.asarray(gdifficults).astype(bool)
    n_gb = int(np.count_nonzero((glabels == label) & ~gdiff))
    gmatch = np.zeros(glabels.shape, bool)
    n = scores.shape[0]
    tp = np.zeros(n, bool)
    fp = np.zeros(n, bool)
    for i in range(n):
        if glabels.shape[0] == 0:
            fp[i] = True
            continue
        jac = bboxes_jaccard(bboxes[i], gbboxes) * (glabels == label).astype(F)
        idx = int(np.argmax(jac))
        match = jac[idx] > F(matching_threshold)
        existing = gmatch[idx]
        nd = not gdiff[idx]
        tp[i] = nd and match and not existing
        fp[i] = nd and (existing or not match)
        if nd and match:
            gmatch[idx] = True
    return n_gb, tp, fp


def voc_ap(rec, prec, use_07_metric):
    """The PASCAL devkit's AP (as in py-faster-rcnn's voc_eval.py), written independently of utility/metrics.py."""
    if use_07_metric:
        ap = 0.
        for t in np.arange(0., 1.1, 0.1):
            p = 0. if np.sum(rec >= t) == 0 else np.max(prec[rec >= t])
            ap += p / 11.
        return ap
    mrec = np.concatenate(([0.], rec, [1.]))
    mpre = np.concatenate(([0.], prec, [0.]))
    for i in range(mpre.size - 1, 0, -1):
        mpre[i - 1] = np.maximum(mpre[i - 1], mpre[i])
    i = np.where(mrec[1:] != mrec[:-1])[0]
    return float(np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1]))
