"""GPU: detection post-processing (xdet_det_postprocess / light_head_rfcn_eval.bboxes_eval) against the numpy
restatement of the reference's eval_helper chain (oracle/detections.py).  Selection is integer work: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import detections as od

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


def make_case(n, r, num_classes, seed, sharp=4.0):
    rng = np.random.default_rng(seed)
    logits = rng.standard_normal((n, r, num_classes)).astype(np.float32) * np.float32(sharp)
    e = np.exp(logits - logits.max(-1, keepdims=True))
    probs = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    # clustered boxes so that NMS has work to do; some degenerate / outside the crop
    ctr = rng.uniform(0.0, 1.0, (n, 24, 2))
    pick = rng.integers(0, 24, (n, r))
    c = np.take_along_axis(ctr, pick[..., None].repeat(2, -1), axis=1) + rng.normal(0, 0.03, (n, r, 2))
    hw = rng.uniform(0.0, 0.5, (n, r, 2))
    boxes = np.concatenate([c - hw / 2, c + hw / 2], -1).astype(np.float32)
    bbox_img = np.stack([np.array([0.0, 0.0, rng.uniform(0.6, 1.0), rng.uniform(0.6, 1.0)], np.float32) for _ in range(n)])
    shapes = np.stack([np.array([rng.integers(200, 600), rng.integers(200, 600)]) for _ in range(n)])
    return probs, boxes, bbox_img, shapes


@pytest.fixture(scope="module")
def lh():
    assert torch.cuda.is_available()
    import xdet_b200  # noqa: F401
    from xdet_b200 import _native
    from xdet_b200 import light_head_rfcn_eval as lh
    _native.lib()
    return lh


@pytest.mark.parametrize("n,r,seed,sharp", [(2, 1000, 0, 4.0), (1, 300, 1, 6.0), (3, 64, 2, 1.0), (1, 1000, 3, 0.2)])
def test_bboxes_eval_matches_oracle(lh, n, r, seed, sharp):
    num_classes = 21
    probs, boxes, bbox_img, shapes = make_case(n, r, num_classes, seed, sharp)
    d_scores, d_boxes = lh.bboxes_eval(shapes, torch.from_numpy(bbox_img).cuda(), torch.from_numpy(probs).cuda(),
                                       torch.from_numpy(boxes).cuda(), num_classes)
    torch.cuda.synchronize()
    assert sorted(d_scores) == list(range(1, num_classes))
    n_det = 0
    for i in range(n):
        ref_s, ref_b = od.bboxes_eval_select(probs[i], boxes[i], bbox_img[i], shapes[i], num_classes)
        for c in range(1, num_classes):
            assert d_scores[c].shape == (n, 200) and d_boxes[c].shape == (n, 200, 4)
            assert np.array_equal(bits(d_scores[c][i].cpu().numpy()), bits(ref_s[c])), (i, c)
            assert np.array_equal(bits(d_boxes[c][i].cpu().numpy()), bits(ref_b[c])), (i, c)
            n_det += int((ref_s[c] > 0).sum())
    if sharp >= 1.0:
        assert n_det > 0  # the case exercises real selections, not only padding


def test_empty_and_saturated_classes(lh):
    """No box above the threshold (all-zero outputs) and more than 2*nms_topk candidates in one class."""
    num_classes = 3
    r = 1200
    rng = np.random.default_rng(7)
    probs = np.zeros((1, r, num_classes), np.float32)
    probs[..., 0] = 0.005
    probs[..., 1] = 0.99 - rng.uniform(0, 0.2, (1, r)).astype(np.float32)   # every box is a class-1 candidate
    probs[..., 2] = 0.005                                                    # class 2: nothing passes 0.01
    c = rng.uniform(0.1, 0.9, (1, r, 2))
    boxes = np.concatenate([c - 0.02, c + 0.02], -1).astype(np.float32)     # small boxes: most survive NMS
    bbox_img = np.array([[0, 0, 1, 1]], np.float32)
    shapes = np.array([[480, 480]])
    d_scores, d_boxes = lh.bboxes_eval(shapes, torch.from_numpy(bbox_img).cuda(), torch.from_numpy(probs).cuda(),
                                       torch.from_numpy(boxes).cuda(), num_classes)
    ref_s, ref_b = od.bboxes_eval_select(probs[0], boxes[0], bbox_img[0], shapes[0], num_classes)
    for c in (1, 2):
        assert np.array_equal(bits(d_scores[c][0].cpu().numpy()), bits(ref_s[c]))
        assert np.array_equal(bits(d_boxes[c][0].cpu().numpy()), bits(ref_b[c]))
    assert not d_scores[2].any() and (d_scores[1][0] > 0).sum() == 200


def test_model_call_with_detections(lh):
    """LightHeadRFCN(..., detections=True): the per-class detections of the model's own head outputs equal the
    oracle chain applied to those outputs (bit-exact selection), also when replayed from a CUDA graph."""
    params = lh.make_params(train_image_size=160, rpn_pre_nms_top_n=300, rpn_post_nms_top_n=64, rpn_min_size=16.0 / 160,
                            select_threshold=0.02)
    model = lh.LightHeadRFCN(params, seed=1)
    g = torch.Generator(device="cuda").manual_seed(4)
    images = torch.rand((2, 3, 160, 160), generator=g, device="cuda") * 2 - 1
    out = model(images, detections=True)
    torch.cuda.synchronize()
    assert out["det_scores"].shape == (2, 20, 200) and out["det_bboxes"].shape == (2, 20, 200, 4)
    probs = out["head_cls_score"].reshape(2, -1, 21).cpu().numpy()
    boxes = out["bboxes_predict"].reshape(2, -1, 4).cpu().numpy()
    for i in range(2):
        ref_s, ref_b = od.bboxes_eval_select(probs[i], boxes[i], np.array([0, 0, 1, 1], np.float32), (160, 160), 21,
                                             select_threshold=0.02, train_image_size=160)
        for c in range(1, 21):
            assert np.array_equal(bits(out["det_scores"][i, c - 1].cpu().numpy()), bits(ref_s[c]))
            assert np.array_equal(bits(out["det_bboxes"][i, c - 1].cpu().numpy()), bits(ref_b[c]))
    # graph replay gives the same detections
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        model(images, detections=True)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out2 = model(images, detections=True)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out2["det_scores"], out["det_scores"]) and torch.equal(out2["det_bboxes"], out["det_bboxes"])


def test_tp_fp_matching_matches_oracle(lh):
    """xdet_det_match (utility.eval_helper.bboxes_matching_batch) vs the loop restatement of bboxes_matching, then the
    whole detections -> TP/FP -> VOC AP chain on the same inputs."""
    from oracle import voc_eval as ov
    from xdet_b200.utility import eval_helper as eh
    from xdet_b200.utility import metrics as M
    num_classes = 21
    n, r = 2, 600
    probs, boxes, bbox_img, shapes = make_case(n, r, num_classes, seed=11, sharp=4.0)
    bbox_img[:] = np.array([0, 0, 1, 1], np.float32)
    rng = np.random.default_rng(12)
    G = 9
    glabels = rng.integers(0, num_classes, (n, G)).astype(np.int64)   # 0 = padding
    gdiff = (rng.random((n, G)) < 0.2).astype(np.int64)
    # ground truth built from a few of the predicted boxes (so that there are true positives), jittered
    pick = rng.integers(0, r, (n, G))
    gboxes = np.clip(np.take_along_axis(boxes, pick[..., None].repeat(4, -1), axis=1) +
                     rng.normal(0, 0.01, (n, G, 4)).astype(np.float32), 0, 1).astype(np.float32)
    d_scores, d_boxes = lh.bboxes_eval(shapes, torch.from_numpy(bbox_img).cuda(), torch.from_numpy(probs).cuda(),
                                       torch.from_numpy(boxes).cuda(), num_classes)
    ngb, tp, fp = eh.bboxes_matching_batch(d_scores.keys(), d_scores, d_boxes, torch.from_numpy(glabels),
                                           torch.from_numpy(gboxes), torch.from_numpy(gdiff))
    torch.cuda.synchronize()
    n_tp = 0
    for c in range(1, num_classes):
        for i in range(n):
            ref_n, ref_tp, ref_fp = ov.bboxes_matching(c, d_scores[c][i].cpu().numpy(), d_boxes[c][i].cpu().numpy(),
                                                       glabels[i], gboxes[i], gdiff[i])
            assert int(ngb[c][i]) == ref_n
            assert np.array_equal(tp[c][i].cpu().numpy(), ref_tp), (c, i)
            assert np.array_equal(fp[c][i].cpu().numpy(), ref_fp), (c, i)
            n_tp += int(ref_tp.sum())
    assert n_tp > 0
    state = M.streaming_tp_fp_arrays(ngb, tp, fp, d_scores)
    m07, aps = M.voc_map(state, use_07_metric=True)
    m12, _ = M.voc_map(state, use_07_metric=False)
    assert 0.0 <= m07 <= 1.0 and 0.0 <= m12 <= 1.0 and len(aps) == num_classes - 1
    # no ground truth at all: every detection is a false positive, none a true positive
    _, tp0, fp0 = eh.bboxes_matching_batch(d_scores.keys(), d_scores, d_boxes, torch.zeros((n, 0), dtype=torch.int64),
                                           torch.zeros((n, 0, 4)), torch.zeros((n, 0), dtype=torch.int64))
    assert not any(bool(tp0[c].any()) for c in tp0) and all(bool(fp0[c].all()) for c in fp0)
