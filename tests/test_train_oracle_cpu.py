"""CPU: the training-step oracle (oracle/net_train.py, autograd over the fp32 restatement) runs for both backbones on a
tiny problem with hand-built selections: finite losses, a gradient for every variable, batch-norm in training mode.
(The GPU parity test of the training step injects the product's own selections; this one keeps the oracle itself
exercised on the CPU, including the XceptionBody variant the product's trainer does not cover yet.)"""
import numpy as np
import pytest
import torch

from oracle import net as onet
from oracle import net_train as ont
from oracle import proposals as op
from oracle import train as ot


@pytest.mark.parametrize("backbone", ["resnet50", "xception"])
def test_oracle_train_step(backbone):
    torch.manual_seed(0)
    size = 96
    params = {"model_scope": "xception_lighthead", "num_classes": 21, "backbone": backbone,
              "rpn_pre_nms_top_n": 100, "rpn_post_nms_top_n": 20, "rpn_nms_thres": 0.7, "rpn_min_size": 16.0 / size,
              "rpn_match_threshold": 0.7, "rpn_neg_threshold": 0.3, "rpn_fg_ratio": 0.5, "match_threshold": 0.5,
              "neg_threshold_high": 0.5, "fg_ratio": 0.25}
    rng = np.random.default_rng(0)
    images = (rng.random((2, 3, size, size), dtype=np.float32) * 2 - 1).astype(np.float32)
    sd = {}
    with torch.no_grad():
        probe = onet.model(images[:1], sd, params, op.layer_anchors((size, size), (6, 6), [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8],
                                                                    [0.1], [1., 2., .5], 16), create_seed=1)
    fm = probe["rpn_feat_map"].shape[2]
    assert fm == 6
    anchors = op.layer_anchors((size, size), (fm, fm), [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5], 16)
    gt_boxes = np.array([[[0.1, 0.1, 0.6, 0.7], [0.5, 0.4, 0.95, 0.9]], [[0.2, 0.3, 0.8, 0.8], [0, 0, 0, 0]]], np.float32)
    gt_labels = np.array([[3, 7], [12, 0]], np.int32)
    A_tot = fm * fm * 22
    # RPN samples: every anchor of the flattened batch (the matching forces at least one positive per ground-truth box)
    rpn_idx = np.arange(0, 2 * A_tot)
    # candidate RoIs: the ground truth, jittered copies and random boxes; 8 sampled per image, OHEM keeps 4
    rois_all = np.zeros((2, 12, 4), np.float32)
    for n in range(2):
        for j in range(12):
            g = gt_boxes[n, j % 2] if gt_labels[n, j % 2] > 0 else gt_boxes[n, 0]
            jit = rng.normal(0, 0.04 * (j // 2), 4).astype(np.float32)
            b = np.clip(g + jit, 0, 1)
            rois_all[n, j] = [min(b[0], b[2] - 0.1), min(b[1], b[3] - 0.1), b[2], b[3]]
    inject = {"rpn_idx": rpn_idx, "rois_all": rois_all, "roi_idx": np.tile(np.arange(8), (2, 1)),
              "ohem_idx": np.array([[0, 2, 5, 7], [1, 3, 4, 6]])}
    losses, grads, mid = ont.train_step(images, gt_boxes, gt_labels, sd, params, anchors, inject)
    assert all(np.isfinite(v) and v > 0 for v in losses.values()), losses
    assert set(grads) <= set(sd) and len(grads) > 40
    nonzero = [k for k, g in grads.items() if float(g.abs().sum()) > 0]
    if backbone == "xception":
        assert sum(1 for k in nonzero if k.endswith("depthwise_kernel")) == 34
        assert any(k.endswith("block14_sepconv2/pointwise_kernel") for k in nonzero)
    else:
        assert any(k.endswith("conv2d/kernel") for k in nonzero)
    # batch statistics were used: the oracle's moving averages are not part of the training graph
    assert not any(k.endswith("moving_mean") or k.endswith("moving_variance") for k in nonzero)
    assert mid["roi_labels"].shape == (2, 8) and (mid["roi_labels"] > 0).any()
    assert ot is not None
