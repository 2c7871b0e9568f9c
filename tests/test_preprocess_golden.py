"""The eval input pipeline against goldens from the reference's own common_preprocessing.py run under the numpy
TensorFlow stand-in (tests/golden/make_preprocess_golden.py).  The golden's bilinear resize is float64 in matrix
form, the oracle / kernel are fp32 lerp by lerp: agreement is to fp32 rounding (2e-6 on values in [-1, 1.2])."""
import os

import numpy as np
import pytest

from oracle import preprocess as op

GOLD = os.path.join(os.path.dirname(__file__), "golden", "preprocess_golden.npz")
TOL = 2e-6


@pytest.fixture(scope="module")
def G():
    return np.load(GOLD)


def test_oracle_matches_reference_preprocessing(G):
    for i in range(4):
        size = tuple(int(v) for v in G["size_%d" % i])
        got = op.preprocess_for_test(G["img_%d" % i], size)
        assert got.shape == G["test_%d" % i].shape == (3,) + size
        assert np.abs(got - G["test_%d" % i]).max() < TOL
    # the eval variant: same image; difficult boxes dropped; bbox_img = the whole image under WARP_RESIZE
    assert np.array_equal(G["eval_image"], G["test_0"])
    keep = G["eval_difficults"] == 0
    assert np.array_equal(G["eval_labels"], G["eval_labels_in"][keep])
    assert np.array_equal(G["eval_bboxes"], G["eval_bboxes_in"][keep])
    assert G["eval_bbox_img"].tolist() == [0.0, 0.0, 1.0, 1.0]


@pytest.mark.gpu
def test_kernel_matches_reference_preprocessing(G):
    import torch
    import xdet_b200  # noqa: F401
    from xdet_b200.preprocessing import common_preprocessing as cp
    for i in range(4):
        size = tuple(int(v) for v in G["size_%d" % i])
        got = cp.light_head_preprocess_for_test(torch.from_numpy(G["img_%d" % i]).cuda(), size)
        torch.cuda.synchronize()
        assert np.abs(got.cpu().numpy() - G["test_%d" % i]).max() < TOL
    img, lab, bb, bbox_img = cp.light_head_preprocess_for_eval(
        torch.from_numpy(G["img_0"]).cuda(), torch.from_numpy(G["eval_labels_in"]).cuda(),
        torch.from_numpy(G["eval_bboxes_in"]).cuda(), tuple(int(v) for v in G["size_0"]),
        difficults=torch.from_numpy(G["eval_difficults"]).cuda())
    assert np.abs(img.cpu().numpy() - G["eval_image"]).max() < TOL
    assert np.array_equal(lab.cpu().numpy(), G["eval_labels"]) and np.array_equal(bb.cpu().numpy(), G["eval_bboxes"])
    assert [float(v) for v in bbox_img.tolist()] == G["eval_bbox_img"].tolist()
