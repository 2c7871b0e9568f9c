"""TP / FP matching of detections against ground truth -- the API of the reference's ``utility/eval_helper.py``
(``bboxes_matching_batch`` :790-840, ``bboxes_matching`` :700-788, ``bboxes_jaccard`` :671-699).

The reference runs one ``tf.while_loop`` per class on the CPU; here all classes and images are one kernel launch
(``xdet_det_match``: one warp per (image, class), detections visited in score order).  The per-class selection that
precedes it (``tf_bboxes_select`` ... ``bboxes_nms_batch``) is ``light_head_rfcn_eval.bboxes_eval``.
"""
import torch

from .. import _native


def _st():
    return torch.cuda.current_stream().cuda_stream


def det_match(det_bboxes, glabels, gbboxes, gdifficults, matching_threshold=0.5):
    """det_bboxes [N,num_fg,K,4] fp32 (class c at index c-1, sorted by score, zero padded), glabels [N,G] (0 = padding),
    gbboxes [N,G,4], gdifficults [N,G] -> tp, fp [N,num_fg,K] bool, n_gbboxes [N,num_fg] int32."""
    N, num_fg, K, _ = det_bboxes.shape
    G = glabels.shape[1]
    dev = det_bboxes.device
    det_bboxes = det_bboxes.contiguous().float()
    gl = glabels.to(dev, torch.int32).contiguous()
    gd = gdifficults.to(dev, torch.int32).contiguous()
    gb = gbboxes.to(dev, torch.float32).contiguous()
    tp = torch.empty((N, num_fg, K), dtype=torch.uint8, device=dev)
    fp = torch.empty_like(tp)
    ngb = torch.empty((N, num_fg), dtype=torch.int32, device=dev)
    rc = _native.lib().xdet_det_match(det_bboxes.data_ptr(), gl.data_ptr(), gb.data_ptr(), gd.data_ptr(), N, num_fg + 1,
                                      K, G, float(matching_threshold), tp.data_ptr(), fp.data_ptr(), ngb.data_ptr(),
                                      _st())
    _native.check(rc)
    return tp.bool(), fp.bool(), ngb


def bboxes_matching_batch(labels, scores, bboxes, glabels, gbboxes, gdifficults, matching_threshold=0.5, scope=None):
    """Reference signature (utility/eval_helper.py:790-840): ``scores`` / ``bboxes`` are the per-class dictionaries
    ``{c: [N,K]}`` / ``{c: [N,K,4]}`` that ``bboxes_eval`` returns, ``labels`` their keys.
    Returns (n_gbboxes {c: [N]}, tp {c: [N,K]}, fp {c: [N,K]})."""
    labels = sorted(labels)
    num_fg = max(labels)
    N, K = bboxes[labels[0]].shape[:2]
    dev = bboxes[labels[0]].device
    det = torch.zeros((N, num_fg, K, 4), dtype=torch.float32, device=dev)
    for c in labels:
        det[:, c - 1] = bboxes[c]
    tp, fp, ngb = det_match(det, glabels, gbboxes, gdifficults, matching_threshold)
    return ({c: ngb[:, c - 1] for c in labels}, {c: tp[:, c - 1] for c in labels}, {c: fp[:, c - 1] for c in labels})
