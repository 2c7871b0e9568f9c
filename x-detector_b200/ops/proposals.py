"""Host wrappers of the RPN proposal kernels in ``csrc/rpn_proposals.cu``."""
import torch

from .. import _native


def _st():
    return torch.cuda.current_stream().cuda_stream


def rpn_decode(rpn_out, cls_off, box_off, anchors, num_anchors):
    """rpn_out [N,fh,fw,ch] fp32 (class logits at cls_off, box deltas at box_off) -> scores [N,fh*fw*A],
    boxes [N,fh*fw*A,4].  anchors = (yref[fh*fw], xref[fh*fw], href[A], wref[A]) fp32 CUDA tensors."""
    N, fh, fw, ch = rpn_out.shape
    A = num_anchors
    assert rpn_out.dtype == torch.float32 and rpn_out.is_contiguous()
    yref, xref, href, wref = anchors
    scores = torch.empty((N, fh * fw * A), dtype=torch.float32, device=rpn_out.device)
    boxes = torch.empty((N, fh * fw * A, 4), dtype=torch.float32, device=rpn_out.device)
    rc = _native.lib().xdet_rpn_decode(rpn_out.data_ptr(), ch, cls_off, box_off, yref.data_ptr(), xref.data_ptr(),
                                       href.data_ptr(), wref.data_ptr(), N, fh, fw, A, scores.data_ptr(),
                                       boxes.data_ptr(), _st())
    _native.check(rc)
    return scores, boxes


_ws_cache = {}


def rpn_select(scores, boxes, pre_nms_top_n, post_nms_top_n, nms_threshold, min_size, shuffle_keys=None,
               return_debug=False):
    """scores [N,A], boxes [N,A,4] -> rois [N,post,4], rois_yxhw [N,post,4], roi_scores [N,post]."""
    N, A_tot = scores.shape
    assert scores.dtype == torch.float32 and boxes.dtype == torch.float32
    scores, boxes = scores.contiguous(), boxes.contiguous()
    dev = scores.device
    need = _native.lib().xdet_rpn_select_workspace_bytes(N, A_tot, pre_nms_top_n)
    key = (dev, torch.cuda.current_stream().cuda_stream)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    rois = torch.empty((N, post_nms_top_n, 4), dtype=torch.float32, device=dev)
    yxhw = torch.empty_like(rois)
    rscore = torch.empty((N, post_nms_top_n), dtype=torch.float32, device=dev)
    keep_idx = torch.empty((N, post_nms_top_n), dtype=torch.int32, device=dev) if return_debug else None
    rc = _native.lib().xdet_rpn_select(scores.data_ptr(), boxes.data_ptr(), N, A_tot, pre_nms_top_n, post_nms_top_n,
                                       float(nms_threshold), float(min_size),
                                       None if shuffle_keys is None else shuffle_keys.data_ptr(), rois.data_ptr(),
                                       yxhw.data_ptr(), rscore.data_ptr(),
                                       None if keep_idx is None else keep_idx.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _st())
    _native.check(rc)
    if return_debug:
        return rois, yxhw, rscore, keep_idx
    return rois, yxhw, rscore


def head_decode(rois, head_out, cls_off, num_classes, loc_off, with_classes=False):
    """rois [M,4], head_out [M,ch] fp32 -> softmax probs [M,num_classes], decoded boxes [M,4]
    (+ with_classes: arg-max class [M] int64 and its probability [M], the predictions dict's 'classes' /
    'probabilities')."""
    rois, head_out = rois.contiguous(), head_out.contiguous()
    M, ch = head_out.shape
    probs = torch.empty((M, num_classes), dtype=torch.float32, device=head_out.device)
    boxes = torch.empty((M, 4), dtype=torch.float32, device=head_out.device)
    classes = torch.empty((M,), dtype=torch.int64, device=head_out.device) if with_classes else None
    best = torch.empty((M,), dtype=torch.float32, device=head_out.device) if with_classes else None
    rc = _native.lib().xdet_head_decode_ex(rois.data_ptr(), head_out.data_ptr(), ch, cls_off, num_classes, loc_off, M,
                                           probs.data_ptr(), boxes.data_ptr(),
                                           None if classes is None else classes.data_ptr(),
                                           None if best is None else best.data_ptr(), _st())
    _native.check(rc)
    if with_classes:
        return probs, boxes, classes, best
    return probs, boxes


def det_postprocess(probs, boxes, bbox_img, min_size, select_threshold, top_k, keep_top_k, nms_threshold):
    """probs [N,R,num_classes], boxes [N,R,4], bbox_img [N,4], min_size [N] (fp32 CUDA) -> per-class detections
    scores [N,num_classes-1,keep_top_k], boxes [N,num_classes-1,keep_top_k,4] (class c at index c-1), descending,
    zero padded: tf_bboxes_select -> bboxes_clip -> filter_boxes -> bboxes_resize -> bboxes_sort -> bboxes_nms_batch
    of the reference's bboxes_eval (light_head_rfcn_eval.py:263-290) as one batched GPU selection."""
    N, R, C = probs.shape
    assert boxes.shape == (N, R, 4) and bbox_img.shape == (N, 4) and min_size.shape == (N,)
    for t in (probs, boxes, bbox_img, min_size):
        assert t.dtype == torch.float32 and t.is_cuda
    probs, boxes, bbox_img, min_size = probs.contiguous(), boxes.contiguous(), bbox_img.contiguous(), min_size.contiguous()
    dev = probs.device
    need = _native.lib().xdet_det_postprocess_workspace_bytes(N, R, C, top_k)
    key = (dev, torch.cuda.current_stream().cuda_stream, "det")
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        _ws_cache[key] = ws
    out_s = torch.empty((N, C - 1, keep_top_k), dtype=torch.float32, device=dev)
    out_b = torch.empty((N, C - 1, keep_top_k, 4), dtype=torch.float32, device=dev)
    rc = _native.lib().xdet_det_postprocess(probs.data_ptr(), boxes.data_ptr(), bbox_img.data_ptr(), min_size.data_ptr(),
                                            N, R, C, float(select_threshold), top_k, keep_top_k, float(nms_threshold),
                                            out_s.data_ptr(), out_b.data_ptr(), ws.data_ptr(), ws.numel(), _st())
    _native.check(rc)
    return out_s, out_b
