"""CPU oracle (TEST INFRASTRUCTURE ONLY -- imported by tests/, never by the product path) for the detection
post-processing of the reference's ``bboxes_eval`` (light_head_rfcn_eval.py:263-290): a numpy fp32 restatement of the
``utility/eval_helper.py`` functions it chains, one function per reference function, operation order kept.

Pinned by the reference's own Python: utility/eval_helper.py is run unmodified under the numpy TensorFlow stand-in
(oracle/tf_shim) to mint tests/golden/tfpath_golden.npz and tests/test_tfpath_golden.py holds this file (and the CUDA
path) to it bit for bit.  Two TensorFlow-core kernels inside that run are restatements, not TensorFlow's code:
``tf.nn.top_k`` (descending, ties -> lower index) and ``tf.image.non_max_suppression`` (TF r1.6 NonMaxSuppressionV2),
both in oracle/proposals.py and reused here -- those two stay unpinned.
"""
import numpy as np

from . import proposals as P

F = np.float32


def tf_bboxes_select_layer(predictions_layer, localizations_layer, select_threshold, num_classes):
    """utility/eval_helper.py:556-588 -> {c: scores [R]}, {c: bboxes [R,4]} for c in 1..num_classes-1."""
    thr = F(0.0 if select_threshold is None else select_threshold)
    d_scores, d_bboxes = {}, {}
    for c in range(1, num_classes):
        scores = predictions_layer[:, c].astype(F)
        fmask = (scores > thr).astype(F)
        d_scores[c] = (scores * fmask).astype(F)
        d_bboxes[c] = (localizations_layer.astype(F) * fmask[:, None]).astype(F)
    return d_scores, d_bboxes


def bboxes_clip(bbox_ref, bboxes):
    """utility/eval_helper.py:365-404."""
    return P.bboxes_clip(bboxes, ref=bbox_ref)


def min_size_of(min_size_ratio, image_shape, net_input_shape):
    """utility/eval_helper.py:295: tf.maximum(0.0001, ratio * tf.sqrt(float32(h*w) / (net_h*net_w)))."""
    q = F(F(int(image_shape[0]) * int(image_shape[1])) / F(int(net_input_shape[0]) * int(net_input_shape[1])))
    return F(max(F(0.0001), F(F(min_size_ratio) * F(np.sqrt(q)))))


def filter_boxes(scores, bboxes, min_size_ratio, image_shape, net_input_shape, keep_top_k=100):
    """utility/eval_helper.py:278-317 (boolean_mask keeps the order; pad_axis only ever pads)."""
    min_size = min_size_of(min_size_ratio, image_shape, net_input_shape)
    ymin, xmin, ymax, xmax = bboxes[:, 0], bboxes[:, 1], bboxes[:, 2], bboxes[:, 3]
    ws = xmax - xmin
    hs = ymax - ymin
    x_ctr = xmin + ws / F(2)
    y_ctr = ymin + hs / F(2)
    keep = (ws > min_size) & (hs > min_size) & (x_ctr > F(0)) & (y_ctr > F(0)) & (x_ctr < F(1)) & (y_ctr < F(1))
    s, b = scores[keep], bboxes[keep]
    pad = max(keep_top_k - s.shape[0], 0)
    return np.concatenate([s, np.zeros(pad, F)]), np.concatenate([b, np.zeros((pad, 4), F)])


def bboxes_resize(bbox_ref, bboxes):
    """utility/eval_helper.py:423-447."""
    r = np.asarray(bbox_ref, F)
    v = np.array([r[0], r[1], r[0], r[1]], F)
    s = np.array([r[2] - r[0], r[3] - r[1], r[2] - r[0], r[3] - r[1]], F)
    return ((bboxes - v) / s).astype(F)


def bboxes_sort(scores, bboxes, top_k):
    """utility/eval_helper.py:333-361 (dict branch: top_k = min(len, top_k); tf.nn.top_k sorted)."""
    k = min(scores.shape[0], top_k)
    order = np.lexsort((np.arange(scores.shape[0]), -scores.astype(np.float64)))[:k]
    return scores[order], bboxes[order]


def bboxes_nms(scores, bboxes, nms_threshold, keep_top_k):
    """utility/eval_helper.py:449-472."""
    s, b, _ = P.bboxes_nms(scores, bboxes, nms_threshold, keep_top_k)
    return s, b


def bboxes_eval_select(cls_pred_prob, bboxes_pred, bbox_img, image_shape, num_classes, select_threshold=0.01,
                       nms_threshold=0.3, nms_topk=200, train_image_size=480, min_size_ratio=0.03):
    """The '/device:CPU:0' block of bboxes_eval (light_head_rfcn_eval.py:274-288) for ONE image:
    -> {c: scores [nms_topk]}, {c: bboxes [nms_topk, 4]}."""
    d_scores, d_bboxes = tf_bboxes_select_layer(cls_pred_prob, bboxes_pred, select_threshold, num_classes)
    out_s, out_b = {}, {}
    for c in d_scores:
        b = bboxes_clip(bbox_img, d_bboxes[c])
        s, b = filter_boxes(d_scores[c], b, min_size_ratio, image_shape, [train_image_size] * 2)
        b = bboxes_resize(bbox_img, b)
        s, b = bboxes_sort(s, b, nms_topk * 2)
        out_s[c], out_b[c] = bboxes_nms(s, b, nms_threshold, nms_topk)
    return out_s, out_b
