"""GPU: RPN proposal kernels vs the numpy restatement (oracle/proposals.py).
Decode: fp32 within 1e-4 (expf vs numpy exp).  Selection (clip/filter/top-k/NMS/upsample) on identical
inputs: bit-exact."""
import numpy as np
import pytest
import torch

from oracle import proposals as op

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    import xdet_b200  # noqa: F401
    from xdet_b200 import ops
    return ops


def anchors_480(fh=30, img=480):
    return op.layer_anchors((img, img), (fh, fh), [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5], 16)


def test_decode_matches_oracle(ops):
    rng = np.random.default_rng(0)
    N, fh, A = 2, 30, 22
    y, x, h, w = anchors_480(fh)
    rpn = rng.standard_normal((N, fh, fh, 6 * A)).astype(np.float32) * 0.5
    s, b = ops.rpn_decode(torch.from_numpy(rpn).cuda(), 0, 2 * A,
                          tuple(torch.from_numpy(np.ascontiguousarray(a.reshape(-1))).cuda() for a in (y, x, h, w)), A)
    so = op.rpn_objectness(rpn[..., :2 * A])
    bo = op.decode_all_anchors(rpn[..., 2 * A:].reshape(N, -1, 4), (y, x, h, w))
    assert np.abs(s.cpu().numpy() - so).max() < 1e-5
    assert np.abs(b.cpu().numpy() - bo).max() < 1e-4


def synth(rng, N, A_tot, spread=0.25):
    """Tie-free scores and boxes with heavy overlap (so NMS really suppresses) and some invalid ones."""
    cy, cx = rng.uniform(-0.1, 1.1, (N, A_tot)), rng.uniform(-0.1, 1.1, (N, A_tot))
    hh, ww = rng.uniform(0.0, spread * 2, (N, A_tot)), rng.uniform(0.0, spread * 2, (N, A_tot))
    boxes = np.stack([cy - hh / 2, cx - ww / 2, cy + hh / 2, cx + ww / 2], -1).astype(np.float32)
    scores = rng.permutation(N * A_tot).reshape(N, A_tot).astype(np.float32)
    scores = ((scores + 1) / (N * A_tot + 2)).astype(np.float32)  # unique, in (0,1)
    return scores, boxes


@pytest.mark.parametrize("cfg", [(2, 19800, 5000, 1000, 0.7), (1, 19800, 10000, 1800, 0.7), (2, 3000, 500, 300, 0.5),
                                 (1, 700, 1000, 200, 0.7), (1, 55000, 5000, 1000, 0.7)])
def test_select_bit_exact(ops, cfg):
    N, A_tot, pre, post, thr = cfg
    rng = np.random.default_rng(sum(cfg[:4]))
    scores, boxes = synth(rng, N, A_tot)
    keys = rng.uniform(0, 1, (N, post)).astype(np.float32)
    min_size = 16.0 / 480
    rois, yxhw, rs, kidx = ops.rpn_select(torch.from_numpy(scores).cuda(), torch.from_numpy(boxes).cuda(), pre, post,
                                          thr, min_size, torch.from_numpy(keys).cuda(), return_debug=True)
    torch.cuda.synchronize()
    ro, so = op.get_proposals(scores, boxes, pre, post, thr, min_size, shuffle_keys=keys)
    assert np.array_equal(rois.cpu().numpy().view(np.int32), ro.view(np.int32))
    assert np.array_equal(rs.cpu().numpy().view(np.int32), so.view(np.int32))
    assert np.array_equal(yxhw.cpu().numpy().view(np.int32), op.point2center(ro).view(np.int32))
    # the NMS survivor positions themselves (before upsampling)
    for n in range(N):
        b = op.bboxes_clip(boxes[n])
        s, bb, _ = op.filter_and_sort_boxes(scores[n], b, min_size, pre)
        idx = op.non_max_suppression_fast(bb, s, post, thr)
        got = kidx[n].cpu().numpy()
        assert np.array_equal(got[:len(idx)], idx) and (got[len(idx):] == -1).all()


def test_select_few_boxes_upsamples_and_default_box(ops):
    rng = np.random.default_rng(5)
    scores, boxes = synth(rng, 2, 400, spread=0.05)
    boxes[1] = 0.0  # image 1: nothing survives the size filter -> default box
    keys = rng.uniform(0, 1, (2, 300)).astype(np.float32)
    rois, _, rs = ops.rpn_select(torch.from_numpy(scores).cuda(), torch.from_numpy(boxes).cuda(), 1000, 300, 0.3,
                                 16.0 / 480, torch.from_numpy(keys).cuda())
    ro, so = op.get_proposals(scores, boxes, 1000, 300, 0.3, 16.0 / 480, shuffle_keys=keys)
    assert np.array_equal(rois.cpu().numpy().view(np.int32), ro.view(np.int32))
    assert np.array_equal(rs.cpu().numpy().view(np.int32), so.view(np.int32))
    assert np.allclose(ro[1], [0.2, 0.2, 0.8, 0.8])


def test_head_decode(ops):
    rng = np.random.default_rng(9)
    M = 1000
    rois = np.ascontiguousarray(np.sort(rng.uniform(0, 1, (M, 2, 2)), axis=1).reshape(M, 4).astype(np.float32))
    head = rng.standard_normal((M, 32)).astype(np.float32)
    p, b = ops.head_decode(torch.from_numpy(rois).cuda(), torch.from_numpy(head).cuda(), 0, 21, 21)
    z = head[:, :21]
    e = np.exp(z - z.max(1, keepdims=True))
    assert np.abs(p.cpu().numpy() - e / e.sum(1, keepdims=True)).max() < 1e-5
    assert np.abs(b.cpu().numpy() - op.ext_decode_rois(rois, head[:, 21:25])).max() < 1e-4
