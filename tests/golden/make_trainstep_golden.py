#!/usr/bin/env python
"""Mint a golden TRAINING STEP (forward: the sample selections and the losses) from the reference's own
light_head_rfcn_train.lighr_head_model_fn (:277-451), called AS A WHOLE and unmodified under the numpy TensorFlow
stand-in (oracle/tf_shim), mode TRAIN, batch of 2:

  XceptionBody / get_rpn / large_sep_kernel with tf.layers.batch_normalization(training=True), the objectness
  softmax, decode_all_anchors, select_samples (RPN fg/bg sampling, :321-358), the RPN cross-entropy and smooth-L1
  losses, get_proposals' training branch -> AnchorEncoder.ext_encode_rois (RoI targets + fg/bg sampling), get_head
  with OHEM (top-k over the per-RoI losses and the reference's axis-1 gather), head_loss_func, the L2 term.

Stand-ins for what TensorFlow / the input pipeline would supply, nothing else:
  * the custom op: tf.load_op_library -> the reference's COMPILED PsRoIAlign (oracle/_ref/libref_psroi.so);
  * tf.random_shuffle -> order by injected keys (told apart by the tf.cond line and file on the stack);
  * the `labels` dict of input_fn (:226-255): AnchorEncoder.encode_all_anchors targets per image and the three
    closures, written here as input_fn writes them;
  * --run_on_cloud=False (the flag the reference's own usage comment passes; True would shell out to cmake).
Decisions are captured where the reference makes them (every tf.gather is logged) and the loss tensors where it names
them (tf.identity(..., name=...)).  Run in the build container only; the .npz is committed.
    python tests/golden/make_trainstep_golden.py"""
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


class PsRoiOpLibrary(object):
    @staticmethod
    def ps_roi_align(inputs, rois, grid_dim_width, grid_dim_height, pool_method):
        out, idx = psroi.psroi_align_fwd(np.asarray(inputs), np.asarray(rois), grid_dim_width, grid_dim_height,
                                         pool_method, impl="ref")
        return tf.constant(out), tf.constant(idx)


tf.OP_LIBRARIES["libps_roi_align.so"] = PsRoiOpLibrary
tf.app.flags.OVERRIDES.update(run_on_cloud=False)
import light_head_rfcn_train as lt  # noqa: E402  (reference)
from preprocessing import anchor_manipulator  # noqa: E402  (reference)

SIZE, N, G = 129, 2, 3
PARAMS = dict(model_scope="xception_lighthead", num_classes=21, data_format="channels_first", batch_size=N,
              rpn_anchors_per_image=64, rpn_fg_ratio=0.5, rpn_pre_nms_top_n=300, rpn_post_nms_top_n=60,
              rpn_nms_thres=0.7, rpn_min_size=16. / SIZE, using_ohem=True, ohem_roi_one_image=16, fg_ratio=0.25,
              weight_decay=1e-4, learning_rate=1e-3, lr_decay_factors=[1., 0.1], decay_boundaries=[1000],
              end_learning_rate=1e-5, momentum=0.9)
ROI_ONE_IMAGE = 32
THRESHOLDS = dict(rpn_match_threshold=0.7, rpn_neg_threshold=0.3, match_threshold=0.5, neg_threshold_high=0.5,
                  neg_threshold_low=0.)


def cond_lines(path, names):
    src = open(path).read().split("\n")
    return {k: [i + 1 for i, ln in enumerate(src) if ln.strip().startswith(k + " = tf.cond")][0] for k in names}


def main():
    assert "tf_shim" in tf.__file__ and psroi.have_ref() and lt.FLAGS.run_on_cloud is False
    rs = np.random.RandomState(2018)
    images = rs.uniform(-1, 1, (N, 3, SIZE, SIZE)).astype(F)
    _layers.reset_variables()
    fm = ((SIZE - 3) // 2 + 1 - 2 + 7) // 8  # block1 (stride 2, valid; valid) then three 'same' stride-2 pools
    creator = anchor_manipulator.AnchorCreator([SIZE] * 2, layers_shapes=[(fm, fm)],
                                               anchor_scales=[[0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]],
                                               extra_anchor_scales=[[0.1]], anchor_ratios=[[1., 2., .5]],
                                               layer_steps=[16])
    all_anchors, num_anchors_list = creator.get_all_anchors()
    coder = anchor_manipulator.AnchorEncoder(all_anchors, num_classes=PARAMS["num_classes"], allowed_borders=[0.],
                                             positive_threshold=THRESHOLDS["rpn_match_threshold"],
                                             ignore_threshold=THRESHOLDS["rpn_neg_threshold"],
                                             prior_scaling=[1., 1., 1., 1.],
                                             rpn_fg_thres=THRESHOLDS["match_threshold"],
                                             rpn_bg_high_thres=THRESHOLDS["neg_threshold_high"],
                                             rpn_bg_low_thres=THRESHOLDS["neg_threshold_low"])
    # ground truth: G boxes per image, close to anchors so that the RPN has positives
    gt = np.zeros((N, G, 4), F)
    for n in range(N):
        for g in range(G):
            cy, cx = rs.uniform(0.3, 0.7, 2)
            h, w = rs.uniform(0.25, 0.55, 2)
            gt[n, g] = [cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2]
    gt = np.clip(gt, 0, 1).astype(F)
    gl = rs.randint(1, 21, (N, G)).astype(np.int64)
    enc = [coder.encode_all_anchors(tf.constant(gl[n]), tf.constant(gt[n])) for n in range(N)]
    glabels = tf.stack([np.asarray(e[0][0]) for e in enc])
    gtargets = tf.stack([np.asarray(e[1][0]) for e in enc])
    gscores = tf.stack([np.asarray(e[2][0]) for e in enc])
    a_tot = int(np.asarray(glabels).size // N)
    n_rpn = N * PARAMS["rpn_anchors_per_image"]

    keys = dict(rpn_fg=rs.uniform(0, 1, N * a_tot), rpn_bg=rs.uniform(0, 1, N * a_tot), rpn_up=rs.uniform(0, 1, n_rpn),
                prop=rs.uniform(0, 1, (N, PARAMS["rpn_post_nms_top_n"])),
                roi_fg=rs.uniform(0, 1, (N, PARAMS["rpn_post_nms_top_n"] + G)),
                roi_bg=rs.uniform(0, 1, (N, PARAMS["rpn_post_nms_top_n"] + G)),
                roi_up=rs.uniform(0, 1, (N, ROI_ONE_IMAGE)))
    keys = {k: v.astype(F) for k, v in keys.items()}
    names = ("fg_select_indices", "bg_select_indices", "final_keep_indices")
    train_lines = cond_lines("/root/reference/light_head_rfcn_train.py", names)
    roi_lines = cond_lines("/root/reference/preprocessing/anchor_manipulator.py", names)

    def keyfn(mi, n, frames):
        for _, line, filename in frames:
            base = os.path.basename(filename)
            for which, short in zip(names, ("fg", "bg", "up")):
                if base == "light_head_rfcn_train.py" and line == train_lines[which]:
                    return keys["rpn_" + short]
                if base == "anchor_manipulator.py" and line == roi_lines[which]:
                    return keys["roi_" + short][mi]
        assert any(os.path.basename(f[2]) == "xception_body.py" and f[0] == "upsampel_impl" for f in frames), frames
        return keys["prop"][mi]

    tf.SHUFFLE_KEYS = keyfn
    tf.GATHER_LOG = []
    captured = {}

    def rpn_encode_fn(rois):  # input_fn's closure (:252), recording what goes in and what comes out
        captured["rois_in"] = np.asarray(rois)
        res = coder.ext_encode_rois(rois, tf.constant(gl), tf.constant(gt), ROI_ONE_IMAGE, PARAMS["fg_ratio"], 0.1,
                                    head_prior_scaling=[1., 1., 1., 1.])
        captured["roi_out"], captured["roi_targets"], captured["roi_labels"], captured["roi_scores"] = (
            np.asarray(v) for v in res)
        return res

    labels = {"targets": [glabels, gtargets, gscores, tf.constant(np.array([[SIZE, SIZE, 3]] * N, np.int64))],
              "rpn_decode_fn": lambda pred: coder.decode_all_anchors([pred], squeeze_inner=True)[0],
              "head_decode_fn": lambda rois, pred: coder.ext_decode_rois(rois, pred, head_prior_scaling=[1., 1., 1., 1.]),
              "rpn_encode_fn": rpn_encode_fn,
              "num_anchors_list": num_anchors_list}
    with np.errstate(all="ignore"):
        spec = lt.lighr_head_model_fn(tf.constant(images), labels, tf.estimator.ModeKeys.TRAIN, PARAMS)

    log = tf.GATHER_LOG
    rpn_idx = [i for s, i, a in log if tuple(s) == (N * a_tot, 2)]
    ohem = [i for s, i, a in log if len(s) == 3 and s[:2] == (N, ROI_ONE_IMAGE) and s[2] == 490 and a == 1]
    assert len(rpn_idx) == 1 and len(ohem) == 1, (len(rpn_idx), len(ohem))
    out = {"gt_boxes": gt, "gt_labels": gl, "glabels": np.asarray(glabels), "gtargets": np.asarray(gtargets),
           "gscores": np.asarray(gscores), "rpn_idx": rpn_idx[0].astype(np.int64), "ohem_idx": ohem[0].astype(np.int64)}
    out.update(captured)
    out.update({"keys_" + k: v for k, v in keys.items()})
    for name in ("rpn_cross_entropy_loss", "rpn_location_loss", "rpn_loss", "head_cross_entropy_loss",
                 "head_location_loss", "head_loss", "total_loss"):
        out[name] = np.asarray(tf.NAMED[name], np.float64)
    assert float(spec.loss) == float(out["total_loss"])
    # ---- what get_init_fn_for_scaffold (utility/train_helper.py:5-72) restores for this graph under the train
    #      script's default flags: the model_fn above already built one Saver through tf.train.Scaffold(init_fn=...);
    #      a second call with a checkpoint that lacks some tensors exercises ignore_missing_vars -------------------
    assert len(tf.SAVERS) == 1
    flags = lt.FLAGS
    restore_all = {k: v.op.name for k, v in tf.SAVERS[0].items()}
    present = sorted(restore_all)
    tf.CHECKPOINT_TENSORS = set(present[::2])
    lt.train_helper.get_init_fn_for_scaffold(flags)
    restore_half = {k: v.op.name for k, v in tf.SAVERS[1].items()}
    out["restore"] = np.array(json.dumps(dict(
        model_scope=flags.model_scope, checkpoint_model_scope=flags.checkpoint_model_scope,
        checkpoint_exclude_scopes=flags.checkpoint_exclude_scopes, ignore_missing_vars=flags.ignore_missing_vars,
        all=restore_all, checkpoint_tensors=sorted(tf.CHECKPOINT_TENSORS), half=restore_half)))
    print("restore map:", len(restore_all), "variables;", len(restore_half), "with half the checkpoint missing")
    out["meta"] = np.array(json.dumps(dict(
        seed=2018, size=SIZE, feature_map=fm, roi_one_image=ROI_ONE_IMAGE, params=PARAMS, thresholds=THRESHOLDS,
        variables=[[k, list(v.shape)] for k, v in _layers.VARIABLES.items()])))
    for k in ("rpn_cross_entropy_loss", "rpn_location_loss", "head_loss", "total_loss"):
        print(k, float(out[k]))
    print("rpn positives sampled:", int((np.asarray(glabels).reshape(-1)[out["rpn_idx"]] > 0).sum()),
          "roi positives:", (captured["roi_labels"] > 0).sum(axis=1), "ohem:", out["ohem_idx"].shape)
    path = os.path.join(HERE, "trainstep_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
