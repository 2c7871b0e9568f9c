"""CPU: the whole evaluation program -- image -> predictions -> per-class detections -> TP/FP matching -> streaming
metric state -- against a golden minted by the reference's own ``lighr_head_model_fn`` (light_head_rfcn_eval.py:
364-446, bboxes_eval :263-362) run as a whole, unmodified, in EVAL mode over three images under the numpy TensorFlow
stand-in (tests/golden/make_evalstep_golden.py).  On this side: oracle/net.py, oracle/detections.py,
oracle/voc_eval.py and the product's host-side utility/metrics.py (the GPU tests compare the kernels with these)."""
import json
import os

import numpy as np
import pytest
import torch

import xdet_b200  # noqa: F401
from oracle import detections as od
from oracle import net as onet
from oracle import proposals as P
from oracle import voc_eval as ov
from xdet_b200.utility import metrics as M

GOLD = os.path.join(os.path.dirname(__file__), "golden", "evalstep_golden.npz")
SCALES, EXTRA, RATIOS = [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5]


@pytest.fixture(scope="module")
def run():
    G = np.load(GOLD)
    meta = json.loads(str(G["meta"]))
    size, flags = meta["size"], meta["flags"]
    sd = {name: torch.from_numpy(onet.seeded_variable(name, tuple(shape))) for name, shape in meta["variables"]}
    params = dict(meta["params"], backbone="xception")
    state, outs = None, []
    for i in range(meta["images"]):
        image = np.random.RandomState(1000 + i).uniform(-1, 1, (1, 3, size, size)).astype(np.float32)
        fm = ((size - 3) // 2 + 1 - 2 + 7) // 8   # block1: stride-2 and stride-1 'valid' convs, then three 'same' pools
        anchors = P.layer_anchors((size, size), (fm, fm), SCALES, EXTRA, RATIOS, 16)
        out = onet.model(image, sd, params, anchors, shuffle_keys=G["keys_%d" % i])
        ds, db = od.bboxes_eval_select(out["head_cls_score"], out["bboxes_predict"], G["bbox_img_%d" % i],
                                       G["shape_%d" % i], params["num_classes"], flags["select_threshold"],
                                       flags["nms_threshold"], flags["nms_topk"], flags["train_image_size"])
        n, tp, fp = {}, {}, {}
        for c in ds:
            n[c], tp[c], fp[c] = ov.bboxes_matching(c, ds[c], db[c], G["gt_labels_%d" % i], G["gt_boxes_%d" % i],
                                                    G["difficult_%d" % i])
        state = M.streaming_tp_fp_arrays(n, tp, fp, ds, state=state)
        outs.append(out)
    return G, meta, outs, state


def test_predictions(run):
    G, meta, outs, _ = run
    for i, out in enumerate(outs):
        probs = out["head_cls_score"]
        assert np.abs(out["bboxes_predict"] - G["pred_%d_bboxes_predict" % i]).max() < 1e-4
        assert np.abs(probs.max(-1) - G["pred_%d_probabilities" % i]).max() < 1e-5
        top2 = np.sort(probs, -1)[:, -2:]
        sure = top2[:, 1] - top2[:, 0] > 1e-5          # argmax is only defined up to fp32 noise on near-ties
        assert sure.sum() > 0.9 * len(sure)
        assert np.array_equal(probs.argmax(-1)[sure], G["pred_%d_classes" % i][sure])


def test_streaming_metric_state(run):
    G, meta, _, state = run
    assert len(state) == 20
    total_tp = 0
    for c, acc in state.items():
        name = meta["label2name"][str(c)]
        nobj, ndet, tp, fp, scores = acc.value()
        assert nobj == int(G["metric_tp_fp_%s_nobjects" % name]), name
        assert ndet == int(G["metric_tp_fp_%s_ndetections" % name]), name
        assert np.array_equal(tp, G["metric_tp_fp_%s_tp" % name]) and np.array_equal(fp, G["metric_tp_fp_%s_fp" % name])
        assert np.abs(scores - G["metric_tp_fp_%s_scores" % name]).max() < 1e-5
        total_tp += int(tp.sum())
    assert total_tp >= 3
