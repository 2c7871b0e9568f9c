#!/usr/bin/env python
"""Mint golden vectors for the NETWORK GRAPH from the reference's own builders, executed unmodified under the numpy
TensorFlow stand-in (oracle/tf_shim: tf.variable_scope + tf.layers restated from the TF 1.6 docs in numpy float64,
see tensorflow/_layers.py):

  xception   light_head_rfcn_eval.py:383-409 replayed call by call: xception_body.XceptionBody (:236-379), get_rpn
             (:381-400), large_sep_kernel (:450-475), objectness softmax, AnchorEncoder.decode_all_anchors,
             get_proposals (:402-448), get_head (:477-560) with the reference's COMPILED PsRoIAlign op
             (oracle/_ref/libref_psroi.so), softmax, ext_decode_rois.
  resnet50   the same head on the ResNet-50 body SURVEY.md section 8 scopes: resnet_v2.conv2d_fixed_padding /
             block_layer(bottleneck_block) for the stem and stages 1-3 (strides 1,2,2), xdet_body.xdet_block_layer
             (xdet_bottleneck_block, dilation 2) for stage 4 -- the reference's blocks, composed as oracle/net.py and
             the product compose them.
  resnet50 training-mode forward (tf.layers.batch_normalization(training=True): batch statistics), batch of 2.

What this pins: every variable NAME and SHAPE the reference graph creates (the checkpoint contract), the wiring
(which tensor feeds which layer, strides, paddings, dilations, scopes, activation placement), and the numbers end
to end.  Variables are a deterministic function of their name (oracle.net.seeded_variable), so the golden holds
outputs only; large feature maps are stored as a channel subsample plus whole-tensor sums.
Run in the build container only (needs /root/reference and oracle/_ref); the .npz is committed.
    python tests/golden/make_netgraph_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), "/root/reference", ROOT]

import numpy as np  # noqa: E402
import tensorflow as tf  # noqa: E402  (the stand-in)
from tensorflow import _layers  # noqa: E402

from net import resnet_v2, xception_body, xdet_body  # noqa: E402  (reference)
from preprocessing import anchor_manipulator  # noqa: E402  (reference)

from oracle import psroi  # noqa: E402  (ctypes driver of the compiled reference op)

F = np.float32
DF = "channels_first"
PARAMS = dict(num_classes=21, rpn_pre_nms_top_n=600, rpn_post_nms_top_n=100, rpn_nms_thres=0.7)
SUBSAMPLE = {"rpn_feat_map": 4, "backbone_feat": 8, "large_sep_feature": 2}
SUBSAMPLE_SQUARE = {"rpn_feat_map": 16, "backbone_feat": 32, "large_sep_feature": 8}   # the xs / rs runs


def image(seed, n, h, w):
    return np.random.RandomState(seed).uniform(-1, 1, (n, 3, h, w)).astype(F)


def put(out, prefix, name, value):
    v = np.asarray(value)
    step = (SUBSAMPLE_SQUARE if prefix in ("xs", "rs") else SUBSAMPLE).get(name, 1)
    out["%s_%s" % (prefix, name)] = v[:, ::step] if step > 1 else v
    if step > 1:
        out["%s_%s_sums" % (prefix, name)] = np.array([v.sum(dtype=np.float64), np.abs(v).sum(dtype=np.float64)])
        out["%s_%s_shape" % (prefix, name)] = np.array(v.shape)


def resnet50_body(x, is_training):
    x = resnet_v2.conv2d_fixed_padding(inputs=x, filters=64, kernel_size=7, strides=2, data_format=DF)
    x = tf.layers.max_pooling2d(inputs=x, pool_size=3, strides=2, padding='SAME', data_format=DF)
    for i, (filters, blocks, strides) in enumerate(((64, 3, 1), (128, 4, 2), (256, 6, 2))):
        x = resnet_v2.block_layer(inputs=x, filters=filters, block_fn=resnet_v2.bottleneck_block, blocks=blocks,
                                  strides=strides, is_training=is_training, name='block_layer%d' % (i + 1),
                                  data_format=DF)
    rpn_feat_map = resnet_v2.batch_norm_relu(x, is_training, DF)
    x = xdet_body.xdet_block_layer(inputs=x, filters=512, block_fn=xdet_body.xdet_bottleneck_block, blocks=3,
                                   dilation_rate=2, is_training=is_training, name='block_layer4', data_format=DF)
    return rpn_feat_map, resnet_v2.batch_norm_relu(x, is_training, DF)


def run(out, prefix, backbone, scope, seed, h, w):
    _layers.reset_variables()
    features = tf.constant(image(seed, 1, h, w))
    with tf.variable_scope(scope, default_name=None, values=[features], reuse=tf.AUTO_REUSE):
        if backbone == "xception":
            rpn_feat_map, backbone_feat = xception_body.XceptionBody(features, PARAMS['num_classes'], is_training=False,
                                                                     data_format=DF)
        else:
            rpn_feat_map, backbone_feat = resnet50_body(features, False)
        fh, fw = rpn_feat_map.shape[2:]
        creator = anchor_manipulator.AnchorCreator([h, w], layers_shapes=[(fh, fw)],
                                                   anchor_scales=[[0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]],
                                                   extra_anchor_scales=[[0.1]], anchor_ratios=[[1., 2., .5]],
                                                   layer_steps=[16])
        all_anchors, num_anchors_list = creator.get_all_anchors()
        coder = anchor_manipulator.AnchorEncoder(all_anchors, num_classes=PARAMS['num_classes'], allowed_borders=[0.],
                                                 positive_threshold=0.7, ignore_threshold=0.3,
                                                 prior_scaling=[1., 1., 1., 1.])
        rpn_cls_score, rpn_bbox_pred = xception_body.get_rpn(rpn_feat_map, num_anchors_list[0], False, DF, 'rpn_head')
        large_sep_feature = xception_body.large_sep_kernel(backbone_feat, 256, 10 * 7 * 7, False, DF,
                                                           'large_sep_feature')
        rpn_cls_score = tf.transpose(rpn_cls_score, [0, 2, 3, 1])
        rpn_bbox_pred = tf.transpose(rpn_bbox_pred, [0, 2, 3, 1])
        put(out, prefix, "rpn_cls", rpn_cls_score)
        put(out, prefix, "rpn_box", rpn_bbox_pred)
        rpn_cls_score = tf.reshape(rpn_cls_score, [-1, 2])
        rpn_object_score = tf.nn.softmax(rpn_cls_score)[:, -1]
        rpn_object_score = tf.reshape(rpn_object_score, [1, -1])
        rpn_location_pred = tf.reshape(rpn_bbox_pred, [1, -1, 4])
        rpn_bboxes_pred = coder.decode_all_anchors([rpn_location_pred], squeeze_inner=True)[0]
        keys = np.random.RandomState(seed + 1).uniform(0, 1, (1, PARAMS['rpn_post_nms_top_n'])).astype(F)
        tf.SHUFFLE_KEYS = keys
        proposals_bboxes = xception_body.get_proposals(rpn_object_score, rpn_bboxes_pred, None,
                                                       PARAMS['rpn_pre_nms_top_n'], PARAMS['rpn_post_nms_top_n'],
                                                       PARAMS['rpn_nms_thres'], 16. / max(h, w), False, DF)
        cls_score, bboxes_reg = xception_body.get_head(
            large_sep_feature,
            lambda input_, bboxes_, gw_, gh_: tuple(tf.constant(a) for a in psroi.psroi_align_fwd(
                np.asarray(input_), np.asarray(bboxes_), gw_, gh_, "max", impl="ref")),
            7, 7, None, proposals_bboxes, PARAMS['num_classes'], False, False, 0, DF, 'final_head')
        head_bboxes_pred = coder.ext_decode_rois(proposals_bboxes, bboxes_reg, head_prior_scaling=[1., 1., 1., 1.])
        head_cls_score = tf.nn.softmax(tf.reshape(cls_score, [-1, PARAMS['num_classes']]))
        head_bboxes_pred = tf.reshape(head_bboxes_pred, [-1, 4])
    for name, v in (("rpn_feat_map", rpn_feat_map), ("backbone_feat", backbone_feat),
                    ("large_sep_feature", large_sep_feature), ("rpn_object_score", rpn_object_score),
                    ("rpn_bboxes_pred", rpn_bboxes_pred), ("shuffle_keys", keys), ("proposals_bboxes", proposals_bboxes),
                    ("cls_score", tf.reshape(cls_score, [-1, PARAMS['num_classes']])),
                    ("bboxes_reg", tf.reshape(bboxes_reg, [-1, 4])), ("head_cls_score", head_cls_score),
                    ("bboxes_predict", head_bboxes_pred)):
        put(out, prefix, name, v)
    out["%s_meta" % prefix] = np.array(json.dumps(dict(
        backbone=backbone, scope=scope, seed=seed, height=h, width=w, rpn_min_size=16. / max(h, w),
        variables=[[k, list(v.shape)] for k, v in _layers.VARIABLES.items()], **PARAMS)))
    print(prefix, "variables:", len(_layers.VARIABLES), "proposals:", np.asarray(proposals_bboxes).shape)


def run_training_forward(out, prefix, scope, seed, h, w):
    _layers.reset_variables()
    features = tf.constant(image(seed, 2, h, w))
    with tf.variable_scope(scope, default_name=None, values=[features], reuse=tf.AUTO_REUSE):
        rpn_feat_map, backbone_feat = resnet50_body(features, True)
        rpn_cls_score, rpn_bbox_pred = xception_body.get_rpn(rpn_feat_map, 22, True, DF, 'rpn_head')
        large_sep_feature = xception_body.large_sep_kernel(backbone_feat, 256, 10 * 7 * 7, True, DF, 'large_sep_feature')
    for name, v in (("rpn_feat_map", rpn_feat_map), ("backbone_feat", backbone_feat),
                    ("large_sep_feature", large_sep_feature), ("rpn_cls", tf.transpose(rpn_cls_score, [0, 2, 3, 1])),
                    ("rpn_box", tf.transpose(rpn_bbox_pred, [0, 2, 3, 1]))):
        put(out, prefix, name, v)
    out["%s_meta" % prefix] = np.array(json.dumps(dict(
        backbone="resnet50", scope=scope, seed=seed, height=h, width=w,
        variables=[[k, list(v.shape)] for k, v in _layers.VARIABLES.items()])))


def main():
    assert "tf_shim" in tf.__file__ and psroi.have_ref()
    out = {}
    run(out, "xc", "xception", "xception_lighthead", 11, 161, 193)
    run(out, "rn", "resnet50", "resnet_lighthead", 12, 145, 177)
    run_training_forward(out, "rt", "resnet_lighthead", 13, 81, 97)
    # square 160x160 inputs: a size the product's launchers take (train_image_size), for the CUDA parity-mode test
    run(out, "xs", "xception", "xception_lighthead", 14, 160, 160)
    run(out, "rs", "resnet50", "resnet_lighthead", 15, 160, 160)
    path = os.path.join(HERE, "netgraph_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
