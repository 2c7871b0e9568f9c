"""CPU: oracle/net.py (the fp32 restatement every GPU model test is judged against) versus goldens minted by the
REFERENCE'S OWN graph builders run unmodified under the numpy TensorFlow stand-in
(tests/golden/make_netgraph_golden.py -> netgraph_golden.npz).  Pins, for both backbones: the set of variable names
and shapes the reference graph creates, the wiring, and the values from the image to the final boxes / scores.

Tolerances: the golden is numpy float64 rounded to fp32 per layer, the oracle is torch fp32 -- agreement is at fp32
rounding level (1e-5 of each tensor's max magnitude over ~60 layers); the discrete stages (top-k, NMS, up-sampling)
must then pick the same proposals."""
import json
import os

import numpy as np
import pytest

from oracle import net as onet
from oracle import proposals as op

GOLD = os.path.join(os.path.dirname(__file__), "golden", "netgraph_golden.npz")
SCALES, EXTRA, RATIOS = [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5]


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


class Recording(dict):
    def __init__(self, *a):
        dict.__init__(self, *a)
        self.used = set()

    def __getitem__(self, k):
        self.used.add(k)
        return dict.__getitem__(self, k)


def state_dict(meta):
    import torch
    return Recording({name: torch.from_numpy(onet.seeded_variable(name, tuple(shape))) for name, shape in meta["variables"]})


def image(meta, n=1):
    return np.random.RandomState(meta["seed"]).uniform(-1, 1, (n, 3, meta["height"], meta["width"])).astype(np.float32)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def check(gold, prefix, name, value, tol=1e-5):
    g = gold["%s_%s" % (prefix, name)]
    v = np.asarray(value)
    if ("%s_%s_shape" % (prefix, name)) in gold.files:  # stored as a channel subsample + whole-tensor sums
        shape = tuple(gold["%s_%s_shape" % (prefix, name)])
        assert v.shape == shape, (name, v.shape, shape)
        step = -(-shape[1] // g.shape[1])
        while v[:, ::step].shape[1] != g.shape[1]:   # ceil(C / step) is not injective: find the stride that was used
            step += 1
        sums = gold["%s_%s_sums" % (prefix, name)]
        assert abs(v.sum(dtype=np.float64) - sums[0]) <= 1e-5 * sums[1], name
        assert abs(np.abs(v).sum(dtype=np.float64) - sums[1]) <= 1e-5 * sums[1], name
        v = v[:, ::step]
    assert rel(v, g) < tol, (name, rel(v, g))


@pytest.mark.parametrize("prefix", ["xc", "rn", "xs", "rs"])
def test_oracle_graph_matches_reference_builders(gold, prefix):
    meta = json.loads(str(gold["%s_meta" % prefix]))
    sd = state_dict(meta)
    params = dict(model_scope=meta["scope"], backbone=meta["backbone"], num_classes=meta["num_classes"],
                  rpn_pre_nms_top_n=meta["rpn_pre_nms_top_n"], rpn_post_nms_top_n=meta["rpn_post_nms_top_n"],
                  rpn_nms_thres=meta["rpn_nms_thres"], rpn_min_size=meta["rpn_min_size"])
    fh, fw = gold["%s_rpn_cls" % prefix].shape[1:3]
    anchors = op.layer_anchors((meta["height"], meta["width"]), (fh, fw), SCALES, EXTRA, RATIOS, 16)
    # create_seed=None: a variable name the reference graph does not create is a KeyError
    out = onet.model(image(meta), sd, params, anchors, shuffle_keys=gold["%s_shuffle_keys" % prefix])
    # ... and every variable the reference creates is consumed (none dropped from the restated graph)
    assert sd.used == set(sd.keys()), sorted(set(sd.keys()) - sd.used)[:5]
    for name in ("rpn_feat_map", "backbone_feat", "large_sep_feature", "rpn_cls", "rpn_box"):
        check(gold, prefix, name, out[name])
    assert np.abs(out["rpn_object_score"] - gold["%s_rpn_object_score" % prefix]).max() < 1e-5
    assert np.abs(out["rpn_bboxes_pred"] - gold["%s_rpn_bboxes_pred" % prefix]).max() < 1e-4
    # same proposals in the same order (coordinates to fp32 noise), without injecting anything
    assert np.abs(out["proposals_bboxes"] - gold["%s_proposals_bboxes" % prefix]).max() < 1e-5
    check(gold, prefix, "cls_score", out["cls_score"], 2e-5)
    check(gold, prefix, "bboxes_reg", out["bboxes_reg"], 2e-5)
    assert np.abs(out["head_cls_score"] - gold["%s_head_cls_score" % prefix]).max() < 1e-5
    assert np.abs(out["bboxes_predict"] - gold["%s_bboxes_predict" % prefix]).max() < 1e-4


def test_variable_inventory(gold):
    """The checkpoint contract: names TF would give the variables (explicit names in XceptionBody; conv2d, conv2d_1,
    ... / batch_normalization, ... in creation order for the ResNet body and the heads)."""
    xc = dict(map(tuple, map(lambda kv: (kv[0], tuple(kv[1])), json.loads(str(gold["xc_meta"]))["variables"])))
    rn = dict(map(tuple, map(lambda kv: (kv[0], tuple(kv[1])), json.loads(str(gold["rn_meta"]))["variables"])))
    assert xc["xception_lighthead/block1_conv1/kernel"] == (3, 3, 3, 32)
    assert xc["xception_lighthead/conv2d_4/kernel"] == (1, 1, 728, 1024)
    assert xc["xception_lighthead/block14_sepconv2/pointwise_kernel"] == (1, 1, 1536, 2048)
    assert xc["xception_lighthead/rpn_head/conv2d/kernel"] == (3, 3, 728, 512)
    assert xc["xception_lighthead/rpn_head/conv2d_2/bias"] == (88,)
    assert xc["xception_lighthead/large_sep_feature/Branch_1/conv2d_1/kernel"] == (1, 15, 256, 490)
    assert xc["xception_lighthead/large_sep_feature/batch_normalization/gamma"] == (490,)
    assert xc["xception_lighthead/final_head/subnet_fc/kernel"] == (490, 2048)
    assert rn["resnet_lighthead/conv2d/kernel"] == (7, 7, 3, 64)
    assert rn["resnet_lighthead/conv2d_52/kernel"] == (1, 1, 512, 2048)   # 1 stem + 16 blocks * 3 + 4 projections
    assert "resnet_lighthead/conv2d_53/kernel" not in rn
    assert rn["resnet_lighthead/batch_normalization_49/beta"] == (2048,)  # 16 * 3 + the two feature-map norms
    assert rn["resnet_lighthead/rpn_head/conv2d/kernel"] == (3, 3, 1024, 512)
    assert not any(k.endswith("/bias") for k in rn if "/rpn_head/" not in k and "/large_sep_feature/" not in k
                   and "/final_head/" not in k)


def test_training_mode_forward(gold):
    """tf.layers.batch_normalization(training=True) through the ResNet body and both heads (batch statistics over
    as few as 2*6*7 values per channel, hence the wider tolerance)."""
    import torch
    meta = json.loads(str(gold["rt_meta"]))
    sd = state_dict(meta)
    nm = onet.Names(sd)
    nm.push(meta["scope"])
    prev = onet.BN_TRAINING
    onet.BN_TRAINING = True
    try:
        with torch.no_grad():
            rpn_feat, backbone = onet.lighthead_resnet50_body(torch.from_numpy(image(meta, 2)), nm)
            cls, box = onet.get_rpn(rpn_feat, nm, "rpn_head")
            thin = onet.large_sep_kernel(backbone, nm, "large_sep_feature")
    finally:
        onet.BN_TRAINING = prev
    assert sd.used == set(sd.keys())
    for name, v in (("rpn_feat_map", rpn_feat), ("backbone_feat", backbone), ("large_sep_feature", thin),
                    ("rpn_cls", cls.permute(0, 2, 3, 1)), ("rpn_box", box.permute(0, 2, 3, 1))):
        check(gold, "rt", name, v.numpy(), 1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f16x2", "fp32x3"])
@pytest.mark.parametrize("prefix", ["xs", "rs"])
def test_cuda_parity_mode_matches_reference_builders(gold, prefix, precision):
    """The CUDA path in the fp32-accurate precisions, fed the same name-seeded variables, against the reference's own graph
    builders (160x160, both backbones): north_star's 1e-4 on boxes / scores, stage tensors at 1e-4 of their magnitude."""
    import torch
    import xdet_b200  # noqa: F401
    from xdet_b200 import light_head_rfcn_eval as lh
    meta = json.loads(str(gold["%s_meta" % prefix]))
    sd = {name: torch.from_numpy(onet.seeded_variable(name, tuple(shape))) for name, shape in meta["variables"]}
    params = lh.make_params(train_image_size=meta["height"], backbone=meta["backbone"], model_scope=meta["scope"],
                            rpn_pre_nms_top_n=meta["rpn_pre_nms_top_n"], rpn_post_nms_top_n=meta["rpn_post_nms_top_n"],
                            rpn_nms_thres=meta["rpn_nms_thres"], rpn_min_size=meta["rpn_min_size"], precision=precision)
    model = lh.LightHeadRFCN(params, seed=0, state_dict=sd)
    keys = torch.from_numpy(gold["%s_shuffle_keys" % prefix]).cuda()
    out = model(torch.from_numpy(image(meta)).cuda(), shuffle_keys=keys)
    torch.cuda.synchronize()
    check(gold, prefix, "rpn_feat_map", out["rpn_feat_map"].permute(0, 3, 1, 2).cpu().numpy(), 1e-4)
    check(gold, prefix, "backbone_feat", out["backbone_feat"].permute(0, 3, 1, 2).cpu().numpy(), 1e-4)
    check(gold, prefix, "large_sep_feature", out["large_sep_feature"].cpu().numpy(), 1e-4)
    rpn = out["rpn_out"].cpu().numpy()
    check(gold, prefix, "rpn_cls", rpn[..., :44], 1e-4)
    check(gold, prefix, "rpn_box", rpn[..., 44:], 1e-4)
    assert np.abs(out["proposals_bboxes"].cpu().numpy() - gold["%s_proposals_bboxes" % prefix]).max() < 1e-4
    assert np.abs(out["cls_score"].cpu().numpy().reshape(-1, 21) - gold["%s_cls_score" % prefix]).max() < 1e-4 * max(
        1.0, np.abs(gold["%s_cls_score" % prefix]).max())
    assert np.abs(out["bboxes_reg"].cpu().numpy().reshape(-1, 4) - gold["%s_bboxes_reg" % prefix]).max() < 1e-4
    assert np.abs(out["head_cls_score"].cpu().numpy() - gold["%s_head_cls_score" % prefix]).max() < 1e-4
    assert np.abs(out["bboxes_predict"].cpu().numpy() - gold["%s_bboxes_predict" % prefix]).max() < 1e-4
