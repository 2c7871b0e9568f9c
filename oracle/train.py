"""CPU restatement (numpy fp32, op order kept) of the TRAINING-side target assignment of the reference.
TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU arm, never by the product.

Follows preprocessing/anchor_manipulator.py: areas/intersection/iou_matrix :22-46, do_dual_max_match :48-94,
AnchorEncoder.encode_anchor :118-171, ext_encode_rois :337-432 and light_head_rfcn_train.py select_samples
:321-358.  tf.random_shuffle(tf.range(n)) is replaced on both sides by the stable argsort of an injected key
array (keys[:n]).  Pinned by the reference's own Python run under the numpy TensorFlow stand-in (oracle/tf_shim):
encode_all_anchors / ext_encode_rois goldens in tests/golden/tfpath_golden.npz, select_samples inside the whole
training model_fn in tests/golden/trainstep_golden.npz (tests/test_tfpath_golden.py, tests/test_trainstep_golden.py).
One documented deviation: unmatched boxes get +0.0 targets where the reference's mask multiplication gives -0.0."""
import numpy as np

f32 = np.float32


def iou_matrix(gt, boxes):
    """gt [G,4], boxes [A,4] (ymin,xmin,ymax,xmax) -> [G,A]."""
    gt, boxes = gt.astype(f32), boxes.astype(f32)
    iy0 = np.maximum(gt[:, None, 0], boxes[None, :, 0])
    ix0 = np.maximum(gt[:, None, 1], boxes[None, :, 1])
    iy1 = np.minimum(gt[:, None, 2], boxes[None, :, 2])
    ix1 = np.minimum(gt[:, None, 3], boxes[None, :, 3])
    inter = np.maximum(iy1 - iy0, f32(0)) * np.maximum(ix1 - ix0, f32(0))
    ag = (gt[:, 3] - gt[:, 1]) * (gt[:, 2] - gt[:, 0])
    ab = (boxes[:, 3] - boxes[:, 1]) * (boxes[:, 2] - boxes[:, 0])
    union = (ag[:, None] + ab[None, :]) - inter
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(union == 0, f32(0), inter / union).astype(f32)


def do_dual_max_match(ov, high, low):
    """ov [G,A] -> (match [A]: gt index, -1 negative, -2 ignore; selected_scores [A])   (:48-94, defaults)."""
    G, A = ov.shape
    anchors_to_gt = ov.argmax(axis=0)
    mv = ov.max(axis=0)
    less = mv < low
    between = (mv < high) & (mv >= low)
    match = np.where(less, -1, anchors_to_gt)
    match = np.where(between, -2, match)
    gt_to_anchors = ov.argmax(axis=1)
    left_mask = np.zeros((G, A), dtype=bool)
    left_mask[np.arange(G), gt_to_anchors] = True
    left_scores = ov * left_mask.astype(f32)
    has = left_mask.any(axis=0)
    forced = left_scores.argmax(axis=0)
    sel = np.where(has, forced, anchors_to_gt)
    return np.where(has, forced, match), ov[sel, np.arange(A)]


def match_encode(boxes, gt, gt_labels, allowed_border, high, low, prior_scaling=(1., 1., 1., 1.), ref_yxhw=None):
    """One image.  boxes [A,4] point form; gt [G,4], gt_labels [G] (<= 0 dropped, tf.boolean_mask);
    ref_yxhw [A,4] = anchors' centre form (encode_anchor) or None = point2center(boxes) (ext_encode_rois).
    -> labels [A] int32, targets [A,4] f32, scores [A] f32."""
    boxes = boxes.astype(f32)
    valid = gt_labels > 0
    gt, gl = gt[valid].astype(f32), gt_labels[valid]
    A = boxes.shape[0]
    if gt.shape[0] == 0:
        return np.zeros(A, np.int32), np.zeros((A, 4), f32), np.zeros(A, f32)
    ymin, xmin, ymax, xmax = boxes.T
    b = f32(allowed_border)
    inside = (ymin >= -b) & (xmin >= -b) & (ymax < f32(1) + b) & (xmax < f32(1) + b)
    ov = iou_matrix(gt, boxes) * inside[None].astype(f32)
    match, scores = do_dual_max_match(ov, f32(high), f32(low))
    mask = match > -1
    idx = np.clip(match, 0, None)
    g = gt[idx]
    labels = gl[idx] * mask + (-1) * (match < -1)
    gcy, gcx = (g[:, 2] + g[:, 0]) / f32(2), (g[:, 3] + g[:, 1]) / f32(2)
    gh, gw = g[:, 2] - g[:, 0], g[:, 3] - g[:, 1]
    if ref_yxhw is None:
        h, w = ymax - ymin, xmax - xmin
        yref, xref = ymin + h / f32(2), xmin + w / f32(2)
    else:
        yref, xref, h, w = ref_yxhw.astype(f32).T
    ps = [f32(p) for p in prior_scaling]
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.stack([(gcy - yref) / h / ps[0], (gcx - xref) / w / ps[1], np.log(gh / h) / ps[2],
                      np.log(gw / w) / ps[3]], axis=-1).astype(f32)
    t = np.where(mask[:, None], t, f32(0))
    return labels.astype(np.int32), t.astype(f32), scores.astype(f32)


def _shuffle(keys, n):
    return np.argsort(keys[:n], kind="stable")


def sample_fg_bg(labels, scores, bg_low, exp_fg, total, keys_fg, keys_bg, keys_up):
    """One group (:394-432 / light_head_rfcn_train.py:321-358) -> (indices [total], (n_pos, n_neg, n_keep))."""
    pos = np.nonzero(labels > 0)[0]
    n_pos = len(pos)
    fg = pos if n_pos < exp_fg else pos[_shuffle(keys_fg, n_pos)[:exp_fg]]
    negm = labels == 0
    if scores is not None:
        negm &= scores > bg_low
    neg = np.nonzero(negm)[0]
    n_neg = len(neg)
    exp_bg = total - min(n_pos, exp_fg)
    bg = neg if n_neg < exp_bg else neg[_shuffle(keys_bg, n_neg)[:exp_bg]]
    keep = np.concatenate([fg, bg])
    n_keep = len(keep)
    if n_keep == 0:
        return np.zeros(total, np.int64), (n_pos, n_neg, 0)
    if n_keep < total:
        left = total - n_keep
        sel = np.concatenate([np.tile(np.arange(n_keep), left // n_keep + 1), _shuffle(keys_up, n_keep)[:left % n_keep]])
        keep = keep[sel]
    return keep, (n_pos, n_neg, n_keep)
