"""CPU: the training-step oracle (oracle/net_train.py, oracle/train.py) versus a golden minted by the reference's own
``lighr_head_model_fn`` (light_head_rfcn_train.py:277-451) run as a whole, unmodified, in TRAIN mode under the numpy
TensorFlow stand-in (tests/golden/make_trainstep_golden.py -> trainstep_golden.npz): the sample selections are
compared exactly, the losses to fp32 noise.  The GPU training tests compare the product with this oracle."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import net as onet
from oracle import net_train as otrain
from oracle import proposals as P
from oracle import train as ot

GOLD = os.path.join(os.path.dirname(__file__), "golden", "trainstep_golden.npz")
SCALES, EXTRA, RATIOS = [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5]
F = np.float32


@pytest.fixture(scope="module")
def G():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def meta(G):
    return json.loads(str(G["meta"]))


def setup(meta):
    size, fm = meta["size"], meta["feature_map"]
    anchors = P.layer_anchors((size, size), (fm, fm), SCALES, EXTRA, RATIOS, 16)
    images = np.random.RandomState(meta["seed"]).uniform(-1, 1, (2, 3, size, size)).astype(F)
    sd = {name: torch.from_numpy(onet.seeded_variable(name, tuple(shape))) for name, shape in meta["variables"]}
    params = dict(meta["params"], backbone="xception", **meta["thresholds"])
    return anchors, images, sd, params


def anchor_points(anchors):
    y, x, h, w = anchors
    fm, A = y.shape[0], h.shape[0]
    cy = np.broadcast_to(y[:, :, None], (fm, fm, A)).reshape(-1).astype(F)
    cx = np.broadcast_to(x[:, :, None], (fm, fm, A)).reshape(-1).astype(F)
    hh = np.broadcast_to(h[None, None, :], (fm, fm, A)).reshape(-1).astype(F)
    ww = np.broadcast_to(w[None, None, :], (fm, fm, A)).reshape(-1).astype(F)
    return (np.stack([cy - hh / F(2), cx - ww / F(2), cy + hh / F(2), cx + ww / F(2)], -1),
            np.stack([cy, cx, hh, ww], -1))


def test_rpn_targets_and_sampling(G, meta):
    """encode_all_anchors per image, then select_samples (:321-358) over the whole batch."""
    anchors, _, _, params = setup(meta)
    pts, ref = anchor_points(anchors)
    labels = []
    for n in range(2):
        l0, t0, s0 = ot.match_encode(pts, G["gt_boxes"][n], G["gt_labels"][n], 0.0, params["rpn_match_threshold"],
                                     params["rpn_neg_threshold"], ref_yxhw=ref)
        assert np.array_equal(l0, G["glabels"][n].reshape(-1)) and np.array_equal(t0, G["gtargets"][n].reshape(-1, 4))
        assert np.array_equal(s0, G["gscores"][n].reshape(-1))
        labels.append(l0)
    n_rpn = params["batch_size"] * params["rpn_anchors_per_image"]
    exp_fg = int(np.rint(F(n_rpn) * F(params["rpn_fg_ratio"])))
    idx, (n_pos, n_neg, n_keep) = ot.sample_fg_bg(np.concatenate(labels), None, 0.0, exp_fg, n_rpn, G["keys_rpn_fg"],
                                                  G["keys_rpn_bg"], G["keys_rpn_up"])
    assert 0 < n_pos < exp_fg and n_keep == n_rpn
    assert np.array_equal(idx, G["rpn_idx"])


def roi_selection(G, params, meta):
    N, R = 2, meta["roi_one_image"]
    rois_all = np.concatenate([G["rois_in"], G["gt_boxes"]], 1)
    roi_idx = np.zeros((N, R), np.int64)
    for n in range(N):
        l1, t1, s1 = ot.match_encode(rois_all[n], G["gt_boxes"][n], G["gt_labels"][n], 0.1, params["match_threshold"],
                                     params["neg_threshold_high"])
        roi_idx[n], _ = ot.sample_fg_bg(l1, s1, params["neg_threshold_low"], int(np.rint(F(R) * F(params["fg_ratio"]))),
                                        R, G["keys_roi_fg"][n], G["keys_roi_bg"][n], G["keys_roi_up"][n])
        assert np.array_equal(rois_all[n][roi_idx[n]], G["roi_out"][n])
        assert np.array_equal(l1[roi_idx[n]], G["roi_labels"][n]) and (l1[roi_idx[n]] > 0).sum() >= 3
        assert np.array_equal(t1[roi_idx[n]], G["roi_targets"][n])
    return rois_all, roi_idx


def test_training_forward_proposals(G, meta):
    """Batch-statistics forward -> objectness -> decode -> get_proposals: the RoIs handed to ext_encode_rois."""
    anchors, images, sd, params = setup(meta)
    nm = onet.Names(sd)
    nm.push(params["model_scope"])
    prev, onet.BN_TRAINING = onet.BN_TRAINING, True
    try:
        with torch.no_grad():
            rpn_feat, _ = onet.xception_body(torch.from_numpy(images), nm)
            cls, box = onet.get_rpn(rpn_feat, nm, "rpn_head")
    finally:
        onet.BN_TRAINING = prev
    score = P.rpn_objectness(cls.permute(0, 2, 3, 1).numpy())
    boxes = P.decode_all_anchors(box.permute(0, 2, 3, 1).numpy().reshape(2, -1, 4), anchors)
    rois, _ = P.get_proposals(score, boxes, params["rpn_pre_nms_top_n"], params["rpn_post_nms_top_n"],
                              params["rpn_nms_thres"], params["rpn_min_size"], G["keys_prop"])
    assert rois.shape == G["rois_in"].shape
    # the same boxes in the same order; coordinates to fp32 noise through exp() of random-weight deltas
    assert np.abs(rois - G["rois_in"]).max() < 1e-4


def test_losses(G, meta):
    anchors, images, sd, params = setup(meta)
    rois_all, roi_idx = roi_selection(G, params, meta)
    inject = {"rpn_idx": G["rpn_idx"], "rois_all": rois_all, "roi_idx": roi_idx, "ohem_idx": G["ohem_idx"]}
    losses, grads, inter = otrain.train_step(images, G["gt_boxes"], G["gt_labels"], sd, params, anchors, inject)
    for k in ("rpn_cross_entropy_loss", "rpn_location_loss", "head_loss"):
        assert abs(losses[k] - float(G[k])) < 2e-5 * max(1.0, abs(float(G[k]))), (k, losses[k], float(G[k]))
    assert np.array_equal(inter["roi_labels"], G["roi_labels"])
    # every trainable variable of the reference graph received a gradient slot, under the reference's name
    trainable = {n for n, _ in meta["variables"] if not n.rsplit("/", 1)[-1].startswith("moving_")}
    assert set(grads.keys()) >= trainable
    assert all(np.isfinite(g.numpy()).all() for g in grads.values())
    # the L2 term (:420): 'batch_normalization' / '_bn' variables are not decayed
    l2 = sum(float((sd[n].double() ** 2).sum()) / 2 for n in trainable if "batch_normalization" not in n and "_bn" not in n)
    parts = sum(float(G[k]) for k in ("rpn_cross_entropy_loss", "rpn_location_loss", "head_loss"))
    assert abs(float(G["total_loss"]) - (parts + params["weight_decay"] * l2)) < 1e-4 * float(G["total_loss"])
    assert abs(float(G["rpn_loss"]) - float(G["rpn_cross_entropy_loss"]) - float(G["rpn_location_loss"])) < 1e-5


def test_ohem_selection_rule(G, meta):
    """get_head with OHEM (:493-521): top-k of the per-RoI loss computed with the CURRENT head weights, per image;
    re-derived here from the oracle's forward (no injection of ohem_idx)."""
    anchors, images, sd, params = setup(meta)
    rois_all, roi_idx = roi_selection(G, params, meta)
    N, R, nc = 2, meta["roi_one_image"], params["num_classes"]
    nm = onet.Names(sd)
    nm.push(params["model_scope"])
    prev, onet.BN_TRAINING = onet.BN_TRAINING, True
    try:
        with torch.no_grad():
            _, backbone = onet.xception_body(torch.from_numpy(images), nm)
            thin = onet.large_sep_kernel(backbone, nm, "large_sep_feature")
            from oracle import psroi
            rois = np.stack([rois_all[n][roi_idx[n]] for n in range(N)])
            yxhw = P.point2center(rois.reshape(-1, 4)).reshape(N, R, 4).astype(F)
            pooled, _ = psroi.psroi_align_fwd(thin.numpy(), np.ascontiguousarray(yxhw), 7, 7, "max")
            feat = torch.from_numpy(pooled.reshape(N * R, -1))
            nm.push("final_head")
            h = onet.dense(feat, nm, "subnet_fc", 2048, relu=True)
            cls_score, bbox_reg = onet.dense(h, nm, "fc_cls", nc), onet.dense(h, nm, "fc_loc", 4)
    finally:
        onet.BN_TRAINING = prev
    lab = torch.from_numpy(G["roi_labels"].reshape(-1).astype(np.int64))
    ce = torch.nn.functional.cross_entropy(cls_score, lab, reduction="none")
    loc = otrain.smooth_l1(bbox_reg - torch.from_numpy(G["roi_targets"].reshape(-1, 4))).sum(-1) * (lab > 0).float()
    loss = (ce + loc / params["fg_ratio"]).reshape(N, R).numpy()
    k = min(params["ohem_roi_one_image"], R)
    for n in range(N):
        order = np.lexsort((np.arange(R), -loss[n].astype(np.float64)))[:k]
        assert np.array_equal(order, G["ohem_idx"][n]), n
