"""The numpy oracles of the TF graph code on the path, pinned to golden vectors minted from the REFERENCE'S OWN Python
(preprocessing/anchor_manipulator.py, net/xception_body.py, utility/eval_helper.py) run unmodified under the numpy
TensorFlow stand-in (tests/golden/make_tfpath_golden.py).  CPU part: oracle == golden, bit for bit.  GPU part: the CUDA
kernels (through the C-ABI) == golden."""
import os

import numpy as np
import pytest

from oracle import detections as od
from oracle import proposals as P
from oracle import voc_eval as ov

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "tfpath_golden.npz"))
F = np.float32
SCALES, EXTRA, RATIOS = [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5]


def bits(a):
    return np.ascontiguousarray(a, dtype=F).view(np.int32)


@pytest.mark.parametrize("size,fm", [(480, 30), (800, 50), (160, 10)])
def test_anchors(size, fm):
    y, x, h, w = P.layer_anchors((size, size), (fm, fm), SCALES, EXTRA, RATIOS, 16)
    for k, a in zip("yxhw", (y, x, h, w)):
        assert np.array_equal(bits(a), bits(G["anchors%d_%s" % (size, k)])), k
    assert int(G["anchors%d_num" % size]) == 22


def test_decode_proposals_point2center_ext_decode():
    anchors = P.layer_anchors((160, 160), (10, 10), SCALES, EXTRA, RATIOS, 16)
    boxes = P.decode_all_anchors(G["dec_pred"], anchors)
    assert np.array_equal(bits(boxes), bits(G["dec_boxes"]))
    rois, _ = P.get_proposals(G["prop_score"], G["dec_boxes"], 600, 100, 0.7, 16 / 160., G["prop_keys"])
    assert np.array_equal(bits(rois), bits(G["prop_rois"]))
    assert np.array_equal(bits(P.point2center(G["prop_rois"].reshape(-1, 4))), bits(G["prop_yxhw"].reshape(-1, 4)))
    ext = P.ext_decode_rois(G["prop_rois"].reshape(-1, 4), G["ext_deltas"].reshape(-1, 4))
    assert np.array_equal(bits(ext), bits(G["ext_boxes"].reshape(-1, 4)))


def test_detection_chain_and_matching():
    s, b = od.bboxes_eval_select(G["det_probs"], G["det_boxes"], G["det_bbox_img"], G["det_shape"], 21)
    for c in range(1, 21):
        assert np.array_equal(bits(s[c]), bits(G["det_out_scores"][c - 1])), c
        assert np.array_equal(bits(b[c]), bits(G["det_out_boxes"][c - 1])), c
        n, tp, fp = ov.bboxes_matching(c, s[c], b[c], G["match_glabels"], G["match_gboxes"], G["match_gdiff"])
        assert n == int(G["match_n"][c - 1])
        assert np.array_equal(tp, G["match_tp"][c - 1]) and np.array_equal(fp, G["match_fp"][c - 1])
    assert int((G["det_out_scores"] > 0).sum()) > 100 and int(G["match_tp"].sum()) > 0


def test_training_targets_and_roi_sampling():
    """encode_all_anchors and ext_encode_rois (matching, target encoding, fg/bg sampling with the injected shuffles).
    Targets of unmatched boxes: the reference multiplies by the 0/1 mask (-0.0 for negative encodings), the oracle
    and the kernels write +0.0 -- equal as numbers, so values are compared numerically and matched rows bit for bit."""
    from oracle import train as ot
    anchors = P.layer_anchors((160, 160), (10, 10), SCALES, EXTRA, RATIOS, 16)
    y, x, h, w = anchors
    fm, A = 10, 22
    cy = np.broadcast_to(y[:, :, None], (fm, fm, A)).reshape(-1).astype(F)
    cx = np.broadcast_to(x[:, :, None], (fm, fm, A)).reshape(-1).astype(F)
    hh = np.broadcast_to(h[None, None, :], (fm, fm, A)).reshape(-1).astype(F)
    ww = np.broadcast_to(w[None, None, :], (fm, fm, A)).reshape(-1).astype(F)
    ref = np.stack([cy, cx, hh, ww], -1)
    pts = np.stack([cy - hh / F(2), cx - ww / F(2), cy + hh / F(2), cx + ww / F(2)], -1)
    gt, gl = G["tgt_gt"], G["tgt_gl"]
    for n in range(2):
        assert np.array_equal(bits(pts), bits(G["enc_points_%d" % n]))
        l0, t0, s0 = ot.match_encode(pts, gt[n], gl[n], 0.0, 0.7, 0.3, ref_yxhw=ref)
        assert np.array_equal(l0, G["enc_labels_%d" % n]) and (l0 > 0).sum() > 0 and (l0 < 0).sum() > 0
        assert np.array_equal(bits(s0), bits(G["enc_scores_%d" % n]))
        assert np.array_equal(t0, G["enc_targets_%d" % n])
        assert np.array_equal(bits(t0[l0 > 0]), bits(G["enc_targets_%d" % n][l0 > 0]))
        # RoI targets + sampling: ground-truth boxes are appended to the RoIs first (ext_encode_rois :352)
        valid = gl[n] > 0
        allr = np.concatenate([G["roi_in"][n], gt[n][valid]], 0)
        l1, t1, s1 = ot.match_encode(allr, gt[n], gl[n], 0.1, 0.5, 0.5)
        keep, _ = ot.sample_fg_bg(l1, s1, 0.0, int(np.rint(F(16) * F(0.25))), 16, G["roi_kfg"][n], G["roi_kbg"][n], G["roi_kup"][n])
        assert np.array_equal(bits(allr[keep]), bits(G["roi_out"][n]))
        assert np.array_equal(l1[keep], G["roi_labels"][n]) and (l1[keep] > 0).sum() == 4
        assert np.array_equal(bits(s1[keep]), bits(G["roi_scores"][n]))
        assert np.array_equal(t1[keep], G["roi_targets"][n])


# ---- the CUDA path against the same goldens ---------------------------------------------------------------------
@pytest.fixture(scope="module")
def cuda():
    import torch
    assert torch.cuda.is_available()
    import xdet_b200  # noqa: F401
    from xdet_b200 import _native
    _native.lib()
    return torch


@pytest.mark.gpu
def test_gpu_anchor_creator(cuda):
    from xdet_b200.preprocessing import anchor_manipulator as am
    for size, fm in ((480, 30), (800, 50), (160, 10)):
        cr = am.AnchorCreator([size, size], layers_shapes=[(fm, fm)], anchor_scales=[SCALES], extra_anchor_scales=[EXTRA],
                              anchor_ratios=[RATIOS], layer_steps=[16])
        anchors, num = cr.get_all_anchors()
        assert num[0] == 22
        for k, a in zip("yxhw", anchors[0]):
            assert np.array_equal(bits(np.asarray(a.cpu() if hasattr(a, "cpu") else a)), bits(G["anchors%d_%s" % (size, k)])), k


@pytest.mark.gpu
def test_gpu_proposals(cuda):
    torch = cuda
    from xdet_b200 import ops
    rois, yxhw, _ = ops.rpn_select(torch.from_numpy(G["prop_score"]).cuda(), torch.from_numpy(G["dec_boxes"]).cuda(), 600, 100,
                                   0.7, 16 / 160., torch.from_numpy(G["prop_keys"]).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(bits(rois.cpu().numpy()), bits(G["prop_rois"]))
    assert np.array_equal(bits(yxhw.cpu().numpy()), bits(G["prop_yxhw"]))


@pytest.mark.gpu
def test_gpu_detection_chain_and_matching(cuda):
    torch = cuda
    from xdet_b200 import light_head_rfcn_eval as lh
    from xdet_b200.utility import eval_helper as eh
    d_s, d_b = lh.bboxes_eval([tuple(int(v) for v in G["det_shape"])], torch.from_numpy(G["det_bbox_img"][None]).cuda(),
                              torch.from_numpy(G["det_probs"][None]).cuda(), torch.from_numpy(G["det_boxes"][None]).cuda(), 21)
    ngb, tp, fp = eh.bboxes_matching_batch(d_s.keys(), d_s, d_b, torch.from_numpy(G["match_glabels"][None]),
                                           torch.from_numpy(G["match_gboxes"][None]), torch.from_numpy(G["match_gdiff"][None]))
    torch.cuda.synchronize()
    for c in range(1, 21):
        assert np.array_equal(bits(d_s[c][0].cpu().numpy()), bits(G["det_out_scores"][c - 1])), c
        assert np.array_equal(bits(d_b[c][0].cpu().numpy()), bits(G["det_out_boxes"][c - 1])), c
        assert int(ngb[c][0]) == int(G["match_n"][c - 1])
        assert np.array_equal(tp[c][0].cpu().numpy(), G["match_tp"][c - 1])
        assert np.array_equal(fp[c][0].cpu().numpy(), G["match_fp"][c - 1])


@pytest.mark.gpu
def test_gpu_anchor_encoder_training_targets(cuda):
    """The reference's method names (preprocessing/anchor_manipulator.py:319-335, :337-432) on the CUDA path against
    the goldens its own AnchorEncoder produced."""
    torch = cuda
    from xdet_b200.preprocessing import anchor_manipulator as am
    cr = am.AnchorCreator([160, 160], layers_shapes=[(10, 10)], anchor_scales=[SCALES], extra_anchor_scales=[EXTRA],
                          anchor_ratios=[RATIOS], layer_steps=[16])
    anchors, _ = cr.get_all_anchors()
    enc = am.AnchorEncoder(anchors, num_classes=21, allowed_borders=[0.], positive_threshold=0.7, ignore_threshold=0.3,
                           prior_scaling=[1., 1., 1., 1.], rpn_fg_thres=0.5, rpn_bg_high_thres=0.5, rpn_bg_low_thres=0.)
    gt, gl = torch.from_numpy(G["tgt_gt"]).cuda(), torch.from_numpy(G["tgt_gl"]).cuda()
    labels, targets, scores, points, n_layers = enc.encode_all_anchors(gl, gt)
    assert n_layers == 1
    # labels / scores / selections are exact; the encoded targets go through the device's logf (the reference: TF's
    # CPU log), so they are held to 2e-6 like the kernel's own test (tests/test_train_ops_gpu.py)
    for n in range(2):
        assert np.array_equal(labels[0][n].cpu().numpy(), G["enc_labels_%d" % n])
        assert np.abs(targets[0][n].cpu().numpy() - G["enc_targets_%d" % n]).max() < 2e-6
        assert np.array_equal(bits(scores[0][n].cpu().numpy()), bits(G["enc_scores_%d" % n]))
        assert np.array_equal(bits(points[0].cpu().numpy()), bits(G["enc_points_%d" % n]))
    keys = {"roi_fg": torch.from_numpy(G["roi_kfg"]).cuda(), "roi_bg": torch.from_numpy(G["roi_kbg"]).cuda(),
            "roi_up": torch.from_numpy(G["roi_kup"]).cuda()}
    rois, tgt, lab, sc = enc.ext_encode_rois(torch.from_numpy(G["roi_in"]).cuda(), gl, gt, 16, 0.25, 0.1, keys=keys)
    assert np.array_equal(bits(rois.cpu().numpy()), bits(G["roi_out"]))
    assert np.array_equal(lab.cpu().numpy(), G["roi_labels"])
    assert np.abs(tgt.cpu().numpy() - G["roi_targets"]).max() < 2e-6
    assert np.array_equal(bits(sc.cpu().numpy()), bits(G["roi_scores"]))


def test_anchor_encoder_wrappers_with_stand_in_kernels(monkeypatch):
    """CPU: the composition logic of AnchorEncoder.encode_all_anchors / ext_encode_rois (argument order, the zero-box
    padding of the appended ground truth, key slicing, gathers) with the two kernel wrappers replaced by the numpy
    oracle behind the wrappers' documented contracts -- against the goldens the reference's own AnchorEncoder produced.
    (Their GPU test is staged; the kernels themselves are validated against the same oracle in test_train_ops_gpu.)"""
    import torch
    import xdet_b200  # noqa: F401
    from oracle import train as ot
    from xdet_b200.ops import train as T
    from xdet_b200.preprocessing import anchor_manipulator as am

    def fake_match_encode(boxes, gt, gt_labels, allowed_border, high_thres, low_thres, prior_scaling=(1., 1., 1., 1.),
                          ref_yxhw=None):
        assert gt_labels.dtype == torch.int32 and gt.is_contiguous() and boxes.is_contiguous()
        outs = []
        for n in range(gt_labels.shape[0]):
            b = (boxes if boxes.dim() == 2 else boxes[n]).numpy()
            outs.append(ot.match_encode(b, gt[n].numpy(), gt_labels[n].numpy(), allowed_border, high_thres, low_thres,
                                        prior_scaling, None if ref_yxhw is None else ref_yxhw.numpy()))
        return tuple(torch.from_numpy(np.stack([o[i] for o in outs])) for i in range(3))

    def fake_sample_fg_bg(labels, scores, bg_low, exp_fg, total, keys_fg, keys_bg, keys_up):
        assert keys_fg.shape == labels.shape and keys_up.shape == (labels.shape[0], total)
        rows = [ot.sample_fg_bg(labels[g].numpy(), None if scores is None else scores[g].numpy(), bg_low, exp_fg, total,
                                keys_fg[g].numpy(), keys_bg[g].numpy(), keys_up[g].numpy()) for g in range(labels.shape[0])]
        return (torch.from_numpy(np.stack([r[0] for r in rows]).astype(np.int32)),
                torch.tensor([r[1] for r in rows], dtype=torch.int32))

    monkeypatch.setattr(T, "match_encode", fake_match_encode)
    monkeypatch.setattr(T, "sample_fg_bg", fake_sample_fg_bg)
    cr = am.AnchorCreator([160, 160], layers_shapes=[(10, 10)], anchor_scales=[SCALES], extra_anchor_scales=[EXTRA],
                          anchor_ratios=[RATIOS], layer_steps=[16])
    anchors, _ = cr.get_all_anchors()
    enc = am.AnchorEncoder(anchors, num_classes=21, allowed_borders=[0.], positive_threshold=0.7, ignore_threshold=0.3,
                           prior_scaling=[1., 1., 1., 1.], rpn_fg_thres=0.5, rpn_bg_high_thres=0.5, rpn_bg_low_thres=0.,
                           device="cpu")
    gt, gl = torch.from_numpy(G["tgt_gt"]), torch.from_numpy(G["tgt_gl"])
    labels, targets, scores, points, n_layers = enc.encode_all_anchors(gl, gt)
    assert n_layers == 1
    for n in range(2):
        assert np.array_equal(labels[0][n].numpy(), G["enc_labels_%d" % n])
        assert np.array_equal(targets[0][n].numpy(), G["enc_targets_%d" % n])
        assert np.array_equal(bits(scores[0][n].numpy()), bits(G["enc_scores_%d" % n]))
        assert np.array_equal(bits(points[0].numpy()), bits(G["enc_points_%d" % n]))
    one = enc.encode_all_anchors(gl[0], gt[0])                     # the reference's per-image form
    assert np.array_equal(one[0][0].numpy(), G["enc_labels_0"]) and one[1][0].shape == (2200, 4)
    keys = {"roi_fg": torch.from_numpy(G["roi_kfg"]), "roi_bg": torch.from_numpy(G["roi_kbg"]),
            "roi_up": torch.from_numpy(G["roi_kup"])}
    rois, tgt, lab, sc = enc.ext_encode_rois(torch.from_numpy(G["roi_in"]), gl, gt, 16, 0.25, 0.1, keys=keys)
    assert np.array_equal(bits(rois.numpy()), bits(G["roi_out"]))
    assert np.array_equal(lab.numpy(), G["roi_labels"]) and np.array_equal(tgt.numpy(), G["roi_targets"])
    assert np.array_equal(bits(sc.numpy()), bits(G["roi_scores"]))
