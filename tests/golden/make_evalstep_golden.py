#!/usr/bin/env python
"""Mint a golden EVALUATION run from the reference's own light_head_rfcn_eval.lighr_head_model_fn (:364-446), called
AS A WHOLE and unmodified under the numpy TensorFlow stand-in (oracle/tf_shim), mode EVAL, once per image for three
images (one call = one run of the metric update ops): XceptionBody -> RPN -> proposals -> PsRoIAlign (the reference's
compiled op) -> head -> predictions, then bboxes_eval (:263-362): per-class select / clip / filter / resize / sort /
NMS, bboxes_matching_batch, streaming_tp_fp_arrays, precision_recall, AP07 / AP12 and the mAPs.

Supplied in TensorFlow's / the input pipeline's place: tf.load_op_library -> oracle/_ref/libref_psroi.so,
tf.random_shuffle -> injected keys, the `labels` dict of input_fn (:190-223), --train_image_size (command line).
Run in the build container only; the .npz is committed.
    python tests/golden/make_evalstep_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), "/root/reference", ROOT]

import numpy as np  # noqa: E402
import tensorflow as tf  # noqa: E402  (the stand-in)
from tensorflow import _layers  # noqa: E402

from oracle import psroi  # noqa: E402  (ctypes driver of the compiled reference op)

F = np.float32
SIZE, IMAGES, G = 161, 3, 6


class PsRoiOpLibrary(object):
    @staticmethod
    def ps_roi_align(inputs, rois, grid_dim_width, grid_dim_height, pool_method):
        out, idx = psroi.psroi_align_fwd(np.asarray(inputs), np.asarray(rois), grid_dim_width, grid_dim_height,
                                         pool_method, impl="ref")
        return tf.constant(out), tf.constant(idx)


tf.OP_LIBRARIES["libps_roi_align.so"] = PsRoiOpLibrary
tf.app.flags.OVERRIDES.update(train_image_size=SIZE)
import light_head_rfcn_eval as le  # noqa: E402  (reference)
from preprocessing import anchor_manipulator  # noqa: E402  (reference)

FLAGS = le.FLAGS
PARAMS = dict(model_scope=FLAGS.model_scope, num_classes=FLAGS.num_classes, data_format="channels_first",
              rpn_pre_nms_top_n=400, rpn_post_nms_top_n=80, rpn_nms_thres=FLAGS.rpn_nms_thres,
              rpn_min_size=16. / SIZE, weight_decay=FLAGS.weight_decay)


def main():
    assert "tf_shim" in tf.__file__ and psroi.have_ref() and FLAGS.train_image_size == SIZE
    rs = np.random.RandomState(77)
    _layers.reset_variables()
    fm = ((SIZE - 3) // 2 + 1 - 2 + 7) // 8
    creator = anchor_manipulator.AnchorCreator([SIZE] * 2, layers_shapes=[(fm, fm)],
                                               anchor_scales=[[0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]],
                                               extra_anchor_scales=[[0.1]], anchor_ratios=[[1., 2., .5]],
                                               layer_steps=[16])
    all_anchors, num_anchors_list = creator.get_all_anchors()
    coder = anchor_manipulator.AnchorEncoder(all_anchors, num_classes=FLAGS.num_classes, allowed_borders=[0.],
                                             positive_threshold=FLAGS.rpn_match_threshold,
                                             ignore_threshold=FLAGS.rpn_neg_threshold, prior_scaling=[1., 1., 1., 1.])
    out = {}
    spec = None
    for i in range(IMAGES):
        image = np.random.RandomState(1000 + i).uniform(-1, 1, (1, 3, SIZE, SIZE)).astype(F)   # the test regenerates it
        gt = np.zeros((1, G, 4), F)
        for g in range(G):
            cy, cx = rs.uniform(0.3, 0.7, 2)
            h, w = rs.uniform(0.35, 0.9, 2)
            gt[0, g] = np.clip([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], 0, 1)
        gl = rs.randint(1, 21, (1, G)).astype(np.int64)
        difficult = np.array([[0, 1, 0, 0, 0, 1]], np.int64)
        bbox_img = np.array([[0., 0., 1., 1.]], F)
        shape = np.array([[375, 500, 3]], np.int64)
        org_image = np.zeros((1, 8, 8, 3), np.uint8)
        keys = rs.uniform(0, 1, (1, PARAMS["rpn_post_nms_top_n"])).astype(F)
        tf.SHUFFLE_KEYS = keys
        valid = gl[0] > 0
        enc = coder.encode_all_anchors(tf.constant(gl[0][valid]), tf.constant(gt[0][valid]))
        targets = [tf.expand_dims(enc[0][0], 0), tf.expand_dims(enc[1][0], 0), tf.expand_dims(enc[2][0], 0),
                   tf.constant(gl), tf.constant(gt), tf.constant(bbox_img), tf.constant(difficult),
                   tf.constant(org_image), tf.constant(shape)]
        labels = {"targets": targets,
                  "rpn_decode_fn": lambda pred: coder.decode_all_anchors([pred], squeeze_inner=True)[0],
                  "head_decode_fn": lambda rois, pred: coder.ext_decode_rois(rois, pred, head_prior_scaling=[1., 1., 1., 1.]),
                  "num_anchors_list": num_anchors_list}
        with np.errstate(all="ignore"):
            spec = le.lighr_head_model_fn(tf.constant(image), labels, tf.estimator.ModeKeys.EVAL, PARAMS)
        out.update({"gt_boxes_%d" % i: gt[0], "gt_labels_%d" % i: gl[0],
                    "difficult_%d" % i: difficult[0], "bbox_img_%d" % i: bbox_img[0], "shape_%d" % i: shape[0],
                    "keys_%d" % i: keys})
        for k in ("classes", "probabilities", "bboxes_predict"):
            out["pred_%d_%s" % (i, k)] = np.asarray(spec.predictions[k])
    # metric values after the last update (what the Estimator reports at the end of the evaluation)
    names = []
    for name, (value, _update) in sorted(spec.eval_metric_ops.items()):
        out["metric_" + name] = np.asarray(value)
        names.append(name)
    out["meta"] = np.array(json.dumps(dict(size=SIZE, images=IMAGES, params=PARAMS, metric_names=names,
                                           label2name={str(k): v for k, v in le.label2name_table.items()},
                                           flags=dict(select_threshold=FLAGS.select_threshold,
                                                      nms_threshold=FLAGS.nms_threshold, nms_topk=FLAGS.nms_topk,
                                                      train_image_size=FLAGS.train_image_size),
                                           variables=[[k, list(v.shape)] for k, v in _layers.VARIABLES.items()])))
    print(len(names), "metric tensors;", names[:5])
    nd = {n: int(out["metric_" + n]) for n in names if n.endswith("ndetections")}
    print("detections per class:", nd)
    print("tp per class:", {n: int(np.asarray(out["metric_" + n]).sum()) for n in names if n.endswith("_tp")})
    path = os.path.join(HERE, "evalstep_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
