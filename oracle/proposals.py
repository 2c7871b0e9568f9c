"""TEST INFRASTRUCTURE ONLY -- numpy fp32 restatement of the RPN proposal path.

Pinned by the reference's own Python: preprocessing/anchor_manipulator.py and net/xception_body.py are run
unmodified under the numpy TensorFlow stand-in (oracle/tf_shim) to mint tests/golden/tfpath_golden.npz, and
tests/test_tfpath_golden.py holds this file (and the CUDA path) to it bit for bit.  Not TensorFlow's own code in
that run, hence still unpinned: tf.nn.top_k, tf.image.non_max_suppression (TF-core kernels, restated below from
TF r1.6) and tf.random_shuffle (injected keys).  Each function cites the reference lines it follows; every
operation is done in float32 in the reference's order (e.g. ``ymin + h / 2.`` and not ``(ymin + ymax) / 2``).

  AnchorCreator.get_layer_anchors    preprocessing/anchor_manipulator.py:698-743
  AnchorEncoder.decode_all_anchors   preprocessing/anchor_manipulator.py:641-669  (center2point :111-112)
  rpn score / loc reshaping          light_head_rfcn_eval.py:389-397
  _bboxes_clip                       net/xception_body.py:173-194
  _filter_and_sort_boxes             net/xception_body.py:133-158  (tf.nn.top_k: descending, ties -> lower index)
  _bboxes_nms                        net/xception_body.py:57-67    (tf.image.non_max_suppression, TF r1.6
                                     core/kernels/non_max_suppression_op.cc -- restated from its published
                                     algorithm: greedy over score order, IoU on min/max-normalised corners,
                                     0 if either area <= 0, suppress iff IoU > threshold (strict))
  _upsample_rois                     net/xception_body.py:196-213  (tf.random_shuffle replaced by an injected
                                     key array: "shuffle" = stable argsort of the first n keys)
  _point2center                      net/xception_body.py:215-218
  ext_decode_rois                    preprocessing/anchor_manipulator.py:671-683
"""
import math

import numpy as np

F = np.float32


def layer_anchors(img_shape, layer_shape, anchor_scale, extra_anchor_scale, anchor_ratio, layer_step, offset=0.5):
    """-> y_on_image[h,w], x_on_image[h,w], h_on_image[A], w_on_image[A] (all float32)."""
    x_on_layer, y_on_layer = np.meshgrid(np.arange(layer_shape[1]), np.arange(layer_shape[0]))
    # (tf.cast(y, float32) + offset) * layer_step / img_shape : python scalars are weak -> float32 ops
    y_on_image = (y_on_layer.astype(F) + F(offset)) * F(layer_step) / F(img_shape[0])
    x_on_image = (x_on_layer.astype(F) + F(offset)) * F(layer_step) / F(img_shape[1])
    hs, ws = [], []
    for s in extra_anchor_scale:
        hs.append(s)
        ws.append(s)
    for s in anchor_scale:
        for r in anchor_ratio:
            hs.append(s / math.sqrt(r))  # python double, then cast (tf.constant(..., float32))
            ws.append(s * math.sqrt(r))
    return y_on_image.astype(F), x_on_image.astype(F), np.array(hs, F), np.array(ws, F)


def decode_all_anchors(pred_location, anchors, prior_scaling=(1., 1., 1., 1.)):
    """pred_location [N, h*w*A, 4] (cy,cx,h,w deltas, anchor order (y, x, a) a fastest) ->
    boxes [N, h*w*A, 4] (ymin,xmin,ymax,xmax)."""
    yref, xref, href, wref = anchors
    h, w = yref.shape
    A = href.shape[0]
    loc = pred_location.reshape(-1, h, w, A, 4).astype(F)
    ps = [F(p) for p in prior_scaling]
    pred_h = np.exp(loc[..., 2] * ps[2]).astype(F) * href
    pred_w = np.exp(loc[..., 3] * ps[3]).astype(F) * wref
    pred_cy = loc[..., 0] * ps[0] * href + yref[..., None]
    pred_cx = loc[..., 1] * ps[1] * wref + xref[..., None]
    out = np.stack([pred_cy - pred_h / F(2), pred_cx - pred_w / F(2), pred_cy + pred_h / F(2),
                    pred_cx + pred_w / F(2)], axis=-1).astype(F)
    return out.reshape(-1, h * w * A, 4)


def rpn_objectness(cls_score_nhwc):
    """[N,h,w,2A] logits -> [N, h*w*A] softmax(...)[:, -1] (light_head_rfcn_eval.py:393-396)."""
    n = cls_score_nhwc.shape[0]
    z = cls_score_nhwc.reshape(-1, 2).astype(F)
    m = z.max(axis=1, keepdims=True)
    e = np.exp(z - m).astype(F)
    p = e / e.sum(axis=1, keepdims=True, dtype=F)
    return p[:, 1].reshape(n, -1).astype(F)


def bboxes_clip(bboxes, ref=(0., 0., 1., 1.)):
    ymin = np.maximum(bboxes[:, 0], F(ref[0]))
    xmin = np.maximum(bboxes[:, 1], F(ref[1]))
    ymax = np.minimum(bboxes[:, 2], F(ref[2]))
    xmax = np.minimum(bboxes[:, 3], F(ref[3]))
    ymin = np.minimum(ymin, ymax)
    xmin = np.minimum(xmin, xmax)
    return np.stack([ymin, xmin, ymax, xmax], axis=1).astype(F)


def filter_and_sort_boxes(scores, bboxes, min_size, keep_topk):
    """-> (scores[keep_topk], boxes[keep_topk,4]) zero padded, plus the selected anchor indices."""
    ymin, xmin, ymax, xmax = bboxes[:, 0], bboxes[:, 1], bboxes[:, 2], bboxes[:, 3]
    ws = xmax - xmin
    hs = ymax - ymin
    x_ctr = xmin + ws / F(2)
    y_ctr = ymin + hs / F(2)
    keep = (ws > F(min_size)) & (hs > F(min_size)) & (x_ctr > F(0)) & (y_ctr > F(0)) & (x_ctr < F(1)) & (y_ctr < F(1))
    idx_all = np.nonzero(keep)[0]
    s = scores[idx_all]
    k = min(s.shape[0], keep_topk)
    order = np.lexsort((np.arange(s.shape[0]), -s.astype(np.float64)))[:k]  # descending, ties -> lower index
    sel = idx_all[order]
    out_s = np.zeros((keep_topk,), F)
    out_b = np.zeros((keep_topk, 4), F)
    out_s[:k] = scores[sel]
    out_b[:k] = bboxes[sel]
    return out_s, out_b, sel


def _iou_greater(b, i, j, thr):
    ymin_i, xmin_i = min(b[i, 0], b[i, 2]), min(b[i, 1], b[i, 3])
    ymax_i, xmax_i = max(b[i, 0], b[i, 2]), max(b[i, 1], b[i, 3])
    ymin_j, xmin_j = min(b[j, 0], b[j, 2]), min(b[j, 1], b[j, 3])
    ymax_j, xmax_j = max(b[j, 0], b[j, 2]), max(b[j, 1], b[j, 3])
    area_i = F(F(ymax_i - ymin_i) * F(xmax_i - xmin_i))
    area_j = F(F(ymax_j - ymin_j) * F(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return False
    iy0, ix0 = max(ymin_i, ymin_j), max(xmin_i, xmin_j)
    iy1, ix1 = min(ymax_i, ymax_j), min(xmax_i, xmax_j)
    inter = F(max(F(iy1 - iy0), F(0)) * max(F(ix1 - ix0), F(0)))
    iou = F(inter / F(F(area_i + area_j) - inter))
    return bool(iou > F(thr))


def non_max_suppression(boxes, scores, max_output_size, iou_threshold):
    """TF r1.6 NonMaxSuppressionV2 semantics -> selected indices (into boxes)."""
    n = boxes.shape[0]
    order = np.lexsort((np.arange(n), -scores.astype(np.float64)))  # std::sort is unstable; fixtures are tie-free
    out_size = min(max_output_size, n)
    selected = []
    b = boxes.astype(F)
    for i in range(n):
        if len(selected) >= out_size:
            break
        cand = order[i]
        ok = True
        for j in reversed(selected):
            if _iou_greater(b, cand, j, iou_threshold):
                ok = False
                break
        if ok:
            selected.append(cand)
    return np.array(selected, np.int64)


def non_max_suppression_fast(boxes, scores, max_output_size, iou_threshold):
    """Same selection as non_max_suppression, vectorised per candidate (float32 ops in the same order)."""
    n = boxes.shape[0]
    order = np.lexsort((np.arange(n), -scores.astype(np.float64)))
    b = boxes.astype(F)
    ymin = np.minimum(b[:, 0], b[:, 2])
    xmin = np.minimum(b[:, 1], b[:, 3])
    ymax = np.maximum(b[:, 0], b[:, 2])
    xmax = np.maximum(b[:, 1], b[:, 3])
    area = ((ymax - ymin) * (xmax - xmin)).astype(F)
    out_size = min(max_output_size, n)
    sel = []
    for i in range(n):
        if len(sel) >= out_size:
            break
        c = order[i]
        if sel and area[c] > 0:
            s = np.array(sel)
            ih = np.maximum(np.minimum(ymax[c], ymax[s]) - np.maximum(ymin[c], ymin[s]), F(0))
            iw = np.maximum(np.minimum(xmax[c], xmax[s]) - np.maximum(xmin[c], xmin[s]), F(0))
            inter = (ih * iw).astype(F)
            with np.errstate(divide="ignore", invalid="ignore"):
                iou = inter / ((area[c] + area[s]).astype(F) - inter)
            if np.any((iou > F(iou_threshold)) & (area[s] > 0)):
                continue
        sel.append(c)
    return np.array(sel, np.int64)


def bboxes_nms(scores, bboxes, nms_threshold, keep_top_k):
    idx = non_max_suppression_fast(bboxes, scores, keep_top_k, nms_threshold)
    out_s = np.zeros((keep_top_k,), F)
    out_b = np.zeros((keep_top_k, 4), F)
    out_s[:idx.shape[0]] = scores[idx]
    out_b[:idx.shape[0]] = bboxes[idx]
    return out_s, out_b, idx


def upsample_rois(scores, bboxes, keep_top_k, shuffle_keys=None):
    """Drop paddings (score <= 0); if nothing is left use the default box; if short, tile and add a
    'random' remainder.  tf.random_shuffle(range(n)) := stable argsort of shuffle_keys[:n]."""
    m = scores > F(0)
    bboxes, scores = bboxes[m], scores[m]
    if scores.shape[0] < 1:
        scores = np.array([1.], F)
        bboxes = np.array([[0.2, 0.2, 0.8, 0.8]], F)
    n = scores.shape[0]
    if n >= keep_top_k:
        return scores, bboxes
    left = keep_top_k - n
    keys = np.arange(n, dtype=F) if shuffle_keys is None else np.asarray(shuffle_keys, F)[:n]
    shuffled = np.argsort(keys, kind="stable")
    sel = np.concatenate([np.tile(np.arange(n), left // n + 1), shuffled[:left % n]])
    return scores[sel], bboxes[sel]


def get_proposals(object_score, bboxes_pred, rpn_pre_nms_top_n, rpn_post_nms_top_n, nms_threshold, rpn_min_size,
                  shuffle_keys=None):
    """Inference branch of get_proposals (net/xception_body.py:402-444): [N,A] scores, [N,A,4] boxes ->
    rois [N, post_n, 4] (ymin,xmin,ymax,xmax) and their scores."""
    rois, rscores = [], []
    for n in range(object_score.shape[0]):
        b = bboxes_clip(bboxes_pred[n].astype(F))
        s, b, _ = filter_and_sort_boxes(object_score[n].astype(F), b, rpn_min_size, rpn_pre_nms_top_n)
        s, b, _ = bboxes_nms(s, b, nms_threshold, rpn_post_nms_top_n)
        k = None if shuffle_keys is None else shuffle_keys[n]
        s, b = upsample_rois(s, b, rpn_post_nms_top_n, k)
        rois.append(b)
        rscores.append(s)
    return np.stack(rois).astype(F), np.stack(rscores).astype(F)


def point2center(boxes):
    ymin, xmin, ymax, xmax = boxes[..., 0], boxes[..., 1], boxes[..., 2], boxes[..., 3]
    h, w = ymax - ymin, xmax - xmin
    return np.stack([ymin + h / F(2), xmin + w / F(2), h, w], axis=-1).astype(F)


def ext_decode_rois(rois, pred, head_prior_scaling=(1., 1., 1., 1.)):
    ps = [F(p) for p in head_prior_scaling]
    href, wref = rois[..., 2] - rois[..., 0], rois[..., 3] - rois[..., 1]
    yref, xref = rois[..., 0] + href / F(2), rois[..., 1] + wref / F(2)
    pred_h = np.exp(pred[..., 2] * ps[2]).astype(F) * href
    pred_w = np.exp(pred[..., 3] * ps[3]).astype(F) * wref
    pred_cy = pred[..., 0] * ps[0] * href + yref
    pred_cx = pred[..., 1] * ps[1] * wref + xref
    return np.stack([pred_cy - pred_h / F(2), pred_cx - pred_w / F(2), pred_cy + pred_h / F(2),
                     pred_cx + pred_w / F(2)], axis=-1).astype(F)
