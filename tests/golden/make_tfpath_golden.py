#!/usr/bin/env python
"""Mint golden vectors from the REFERENCE'S OWN Python graph code, executed unmodified under the numpy TensorFlow
stand-in oracle/tf_shim (TensorFlow 1.6 itself cannot be installed offline):

  preprocessing/anchor_manipulator.py  AnchorCreator.get_all_anchors (:698-757), AnchorEncoder.decode_all_anchors
                                       (:641-669), ext_decode_rois (:671-683)
  net/xception_body.py                 get_proposals (:402-448) = _bboxes_clip, _filter_and_sort_boxes, _bboxes_nms,
                                       _upsample_rois; _point2center (:215-218)
                                       AnchorEncoder.encode_all_anchors (:118-171,319-335: iou_matrix, do_dual_max_match,
                                       target encoding) and ext_encode_rois (:337-432: RoI targets + fg/bg sampling)
  utility/eval_helper.py               tf_bboxes_select, bboxes_clip, filter_boxes, bboxes_resize, bboxes_sort,
                                       bboxes_nms_batch (the chain of light_head_rfcn_eval.py:274-288), bboxes_matching

The stand-in supplies tf.nn.top_k, tf.image.non_max_suppression (restated TF r1.6 kernels) and tf.random_shuffle (the
injected-key order); everything around them is the reference's code.  Run in the build container only; the .npz is
committed (the GPU box has no /root/reference).
    python tests/golden/make_tfpath_golden.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), "/root/reference", ROOT]

import numpy as np  # noqa: E402
import tensorflow as tf  # noqa: E402  (the stand-in)

from net import xception_body as xb  # noqa: E402  (reference)
from preprocessing import anchor_manipulator as am  # noqa: E402  (reference)
from utility import eval_helper as eh  # noqa: E402  (reference)

F = np.float32
SCALES, EXTRA, RATIOS = [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5]


def main():
    assert "tf_shim" in tf.__file__
    out = {}
    rng = np.random.default_rng(1806)
    # ---- anchors for the 480 and 800 configurations -----------------------------------------------------------
    for size, fm in ((480, 30), (800, 50), (160, 10)):
        cr = am.AnchorCreator([size, size], layers_shapes=[(fm, fm)], anchor_scales=[SCALES], extra_anchor_scales=[EXTRA],
                              anchor_ratios=[RATIOS], layer_steps=[16])
        anchors, num = cr.get_all_anchors()
        for k, a in zip("yxhw", anchors[0]):
            out["anchors%d_%s" % (size, k)] = np.asarray(a)
        out["anchors%d_num" % size] = np.int64(num[0])
    # ---- decode / proposals on a 160x160 problem (10x10 map, 2200 anchors) -------------------------------------
    cr = am.AnchorCreator([160, 160], layers_shapes=[(10, 10)], anchor_scales=[SCALES], extra_anchor_scales=[EXTRA],
                          anchor_ratios=[RATIOS], layer_steps=[16])
    anchors, num = cr.get_all_anchors()
    enc = am.AnchorEncoder(anchors, num_classes=21, allowed_borders=[0.], positive_threshold=0.7, ignore_threshold=0.3,
                           prior_scaling=[1., 1., 1., 1.])
    pred = (rng.standard_normal((3, 2200, 4)) * 0.5).astype(F)
    boxes = np.asarray(enc.decode_all_anchors([tf.constant(pred)], squeeze_inner=True)[0])
    out["dec_pred"], out["dec_boxes"] = pred, boxes
    score = rng.random((3, 2200)).astype(F)
    score[2, 40:] = 0  # an image with very few candidates: exercises the up-sampling with the injected shuffle
    keys = rng.random((3, 100)).astype(F)
    tf.SHUFFLE_KEYS = keys
    rois = np.asarray(xb.get_proposals(tf.constant(score), tf.constant(boxes), None, 600, 100, 0.7, 16 / 160., False,
                                       'channels_last'))
    out["prop_score"], out["prop_keys"], out["prop_rois"] = score, keys, rois
    out["prop_yxhw"] = np.asarray(xb._point2center(tf.constant(rois)))
    deltas = (rng.standard_normal((3, 100, 4)) * 0.3).astype(F)
    out["ext_deltas"] = deltas
    out["ext_boxes"] = np.asarray(enc.ext_decode_rois(tf.constant(rois), tf.constant(deltas)))
    # ---- detection post-processing + matching for one image ----------------------------------------------------
    nc = 21
    logits = (rng.standard_normal((700, nc)) * 4).astype(F)
    e = np.exp(logits - logits.max(-1, keepdims=True))
    probs = (e / e.sum(-1, keepdims=True)).astype(F)
    ctr = rng.uniform(0, 1, (24, 2))
    c = ctr[rng.integers(0, 24, 700)] + rng.normal(0, 0.03, (700, 2))
    hw = rng.uniform(0, 0.5, (700, 2))
    det = np.concatenate([c - hw / 2, c + hw / 2], -1).astype(F)
    bbox_img = np.array([0.0, 0.0, 0.85, 0.9], F)
    shape = np.array([375, 500], np.int64)
    s, b = eh.tf_bboxes_select([tf.constant(probs)], [tf.constant(det)], 0.01, nc, scope='x')
    b = eh.bboxes_clip(tf.constant(bbox_img), b)
    s, b = eh.filter_boxes(s, b, 0.03, tf.constant(shape), [480] * 2, keep_top_k=400)
    b = eh.bboxes_resize(tf.constant(bbox_img), b)
    s, b = eh.bboxes_sort(s, b, top_k=400)
    s, b = eh.bboxes_nms_batch(s, b, nms_threshold=0.3, keep_top_k=200)
    out["det_probs"], out["det_boxes"], out["det_bbox_img"], out["det_shape"] = probs, det, bbox_img, shape
    out["det_out_scores"] = np.stack([np.asarray(s[k]) for k in range(1, nc)])
    out["det_out_boxes"] = np.stack([np.asarray(b[k]) for k in range(1, nc)])
    # ground truth: jittered copies of some output detections (so that true positives, duplicates = false positives,
    # 'difficult' objects and unmatched objects all occur), plus padding (label 0)
    gl, gb = [], []
    for k in (1, 2, 3, 5, 8, 13, 20):
        for j in (0, 3):
            if float(np.asarray(s[k])[j]) > 0:
                gl.append(k)
                gb.append(np.asarray(b[k])[j] + rng.normal(0, 0.004, 4).astype(F))
    gl += [4, 0, 0]
    gb += [np.array([0.05, 0.05, 0.1, 0.1], F), np.zeros(4, F), np.zeros(4, F)]
    G = len(gl)
    glabels = np.array(gl, np.int64)
    gboxes = np.clip(np.stack(gb), 0, 1).astype(F)
    gdiff = (rng.random(G) < 0.25).astype(np.int64)
    tps, fps, ngs = [], [], []
    for k in range(1, nc):
        n, tp, fp = eh.bboxes_matching(k, s[k], b[k], tf.constant(glabels), tf.constant(gboxes), tf.constant(gdiff))
        ngs.append(int(n))
        tps.append(np.asarray(tp))
        fps.append(np.asarray(fp))
    out["match_glabels"], out["match_gboxes"], out["match_gdiff"] = glabels, gboxes, gdiff
    out["match_n"], out["match_tp"], out["match_fp"] = np.array(ngs, np.int64), np.stack(tps), np.stack(fps)
    # ---- training targets: anchor matching / encoding, RoI targets + sampling ------------------------------------
    enc2 = am.AnchorEncoder(anchors, num_classes=21, allowed_borders=[0.], positive_threshold=0.7, ignore_threshold=0.3,
                            prior_scaling=[1., 1., 1., 1.], rpn_fg_thres=0.5, rpn_bg_high_thres=0.5, rpn_bg_low_thres=0.)
    N, R, Gn = 2, 60, 4
    gt = np.zeros((N, Gn, 4), F)
    gl = np.zeros((N, Gn), np.int64)
    gt[0, :3] = [[0.1, 0.1, 0.6, 0.7], [0.5, 0.4, 0.95, 0.9], [0.3, 0.05, 0.5, 0.3]]
    gl[0, :3] = [3, 7, 12]
    gt[1, :2] = [[0.2, 0.3, 0.8, 0.8], [0.05, 0.5, 0.4, 0.95]]
    gl[1, :2] = [1, 20]
    out["tgt_gt"], out["tgt_gl"] = gt, gl
    with np.errstate(all="ignore"):
        for n in range(N):
            valid = gl[n] > 0
            labels, targets, scores, pts, _ = enc2.encode_all_anchors(tf.constant(gl[n][valid]), tf.constant(gt[n][valid]))
            out["enc_labels_%d" % n] = np.asarray(labels[0]).reshape(-1)
            out["enc_targets_%d" % n] = np.asarray(targets[0]).reshape(-1, 4)
            out["enc_scores_%d" % n] = np.asarray(scores[0]).reshape(-1)
            out["enc_points_%d" % n] = np.asarray(pts[0])
    rois_in = np.zeros((N, R, 4), F)
    for n in range(N):
        ng = int((gl[n] > 0).sum())
        for j in range(R):
            bb = np.clip(gt[n, j % ng] + rng.normal(0, 0.02 * (j // 3), 4).astype(F), 0, 1)
            rois_in[n, j] = [min(bb[0], bb[2] - 0.05), min(bb[1], bb[3] - 0.05), bb[2], bb[3]]
    rois_in = np.clip(rois_in, 0, 1).astype(F)
    per_image, fg_frac = 16, 0.25
    kfg, kbg = rng.random((N, R + Gn)).astype(F), rng.random((N, R + Gn)).astype(F)
    kup = rng.random((N, per_image)).astype(F)
    src = open("/root/reference/preprocessing/anchor_manipulator.py").read().split("\n")
    line_of = {k: [i + 1 for i, ln in enumerate(src) if ln.strip().startswith(k + " = tf.cond")][0]
               for k in ("fg_select_indices", "bg_select_indices", "final_keep_indices")}

    def keyfn(mi, n, frames):  # which of ext_encode_rois' three shuffles is running: by the tf.cond line on the stack
        lines = [f[1] for f in frames]
        if line_of["fg_select_indices"] in lines:
            return kfg[mi]
        if line_of["bg_select_indices"] in lines:
            return kbg[mi]
        assert line_of["final_keep_indices"] in lines
        return kup[mi]

    tf.SHUFFLE_KEYS = keyfn
    with np.errstate(all="ignore"):
        pr, pt, pl, ps = enc2.ext_encode_rois(tf.constant(rois_in), tf.constant(gl), tf.constant(gt), per_image, fg_frac, 0.1)
    out["roi_in"], out["roi_kfg"], out["roi_kbg"], out["roi_kup"] = rois_in, kfg, kbg, kup
    out["roi_out"], out["roi_targets"], out["roi_labels"], out["roi_scores"] = (np.asarray(v) for v in (pr, pt, pl, ps))
    path = os.path.join(HERE, "tfpath_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "%.0f KB" % (os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
