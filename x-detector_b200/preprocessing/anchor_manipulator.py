"""``AnchorCreator`` / ``AnchorEncoder`` -- the slice of the reference's
``preprocessing/anchor_manipulator.py`` that the inference hot path uses, same constructor arguments.

* ``AnchorCreator.get_layer_anchors`` / ``get_all_anchors`` (reference :698-757): anchor centres and sizes as
  fp32 constants, computed once on the host exactly as the reference does (python-double sqrt, then cast).
* ``AnchorEncoder.decode_all_anchors`` (:641-669) and ``ext_decode_rois`` (:671-683): CUDA kernels
  ``xdet_rpn_decode`` / ``xdet_head_decode``.
* ``encode_all_anchors`` / ``ext_encode_rois`` (training targets, :319-636) are NOT part of this build.
"""
import math

import numpy as np
import torch

from .. import ops


class AnchorCreator(object):
    def __init__(self, img_shape, layers_shapes, anchor_scales, extra_anchor_scales, anchor_ratios, layer_steps):
        super(AnchorCreator, self).__init__()
        # img_shape -> (height, width)
        self._img_shape = img_shape
        self._layers_shapes = layers_shapes
        self._anchor_scales = anchor_scales
        self._extra_anchor_scales = extra_anchor_scales
        self._anchor_ratios = anchor_ratios
        self._layer_steps = layer_steps
        self._anchor_offset = [0.5] * len(self._layers_shapes)

    def get_layer_anchors(self, layer_shape, anchor_scale, extra_anchor_scale, anchor_ratio, layer_step, offset=0.5):
        f = np.float32
        x_on_layer, y_on_layer = np.meshgrid(np.arange(layer_shape[1]), np.arange(layer_shape[0]))
        y_on_image = (y_on_layer.astype(f) + f(offset)) * f(layer_step) / f(self._img_shape[0])
        x_on_image = (x_on_layer.astype(f) + f(offset)) * f(layer_step) / f(self._img_shape[1])
        num_anchors = len(anchor_scale) * len(anchor_ratio) + len(extra_anchor_scale)
        list_h_on_image, list_w_on_image = [], []
        for scale in extra_anchor_scale:
            list_h_on_image.append(scale)
            list_w_on_image.append(scale)
        for scale in anchor_scale:
            for ratio in anchor_ratio:
                list_h_on_image.append(scale / math.sqrt(ratio))
                list_w_on_image.append(scale * math.sqrt(ratio))
        return (y_on_image.astype(f), x_on_image.astype(f), np.array(list_h_on_image, f), np.array(list_w_on_image, f),
                num_anchors)

    def get_all_anchors(self):
        all_anchors, num_anchors = [], []
        for layer_index, layer_shape in enumerate(self._layers_shapes):
            a = self.get_layer_anchors(layer_shape, self._anchor_scales[layer_index],
                                       self._extra_anchor_scales[layer_index], self._anchor_ratios[layer_index],
                                       self._layer_steps[layer_index], self._anchor_offset[layer_index])
            all_anchors.append(a[:-1])
            num_anchors.append(a[-1])
        return all_anchors, num_anchors


class AnchorEncoder(object):
    def __init__(self, anchors, num_classes, allowed_borders, positive_threshold, ignore_threshold, prior_scaling,
                 rpn_fg_thres=0.5, rpn_bg_high_thres=0.5, rpn_bg_low_thres=0., device="cuda"):
        super(AnchorEncoder, self).__init__()
        self._anchors = anchors
        self._num_classes = num_classes
        self._allowed_borders = allowed_borders
        self._positive_threshold = positive_threshold
        self._ignore_threshold = ignore_threshold
        self._prior_scaling = prior_scaling
        self._rpn_fg_thres = rpn_fg_thres
        self._rpn_bg_high_thres = rpn_bg_high_thres
        self._rpn_bg_low_thres = rpn_bg_low_thres
        if any(float(p) != 1.0 for p in prior_scaling):
            raise ValueError("only prior_scaling = [1,1,1,1] (the reference's setting, light_head_rfcn_train.py:234) "
                             "is built")
        self._dev_anchors = [tuple(torch.from_numpy(np.ascontiguousarray(a.reshape(-1))).to(device) for a in layer)
                             for layer in anchors]

    def device_anchors(self, index=0):
        return self._dev_anchors[index]

    def center2point(self, center_y, center_x, height, width):
        return center_y - height / 2., center_x - width / 2., center_y + height / 2., center_x + width / 2.,

    def point2center(self, ymin, xmin, ymax, xmax):
        height, width = (ymax - ymin), (xmax - xmin)
        return ymin + height / 2., xmin + width / 2., height, width

    def decode_all_anchors(self, pred_location, squeeze_inner=False):
        """pred_location: list (one per feature layer) of [N, fh*fw*A, 4] (or [N,fh,fw,4A]) fp32 CUDA tensors ->
        list of decoded boxes [N, fh*fw*A, 4] (ymin,xmin,ymax,xmax)."""
        assert len(self._anchors) == len(pred_location), 'predict location not equals to anchor priors.'
        pred_bboxes = []
        for index, location_ in enumerate(pred_location):
            y, x, h, w = self._anchors[index]
            fh, fw, A = y.shape[0], y.shape[1], h.shape[0]
            loc = location_.reshape(-1, fh, fw, 4 * A).contiguous()
            _, boxes = ops.rpn_decode(loc, 0, 0, self._dev_anchors[index], A)  # logits slot unused here
            pred_bboxes.append(boxes if squeeze_inner else boxes.reshape(-1, fh, fw, A, 4))
        return pred_bboxes

    def ext_decode_rois(self, proposals_roi, pred_location, head_prior_scaling=[1., 1., 1., 1.]):
        if any(float(p) != 1.0 for p in head_prior_scaling):
            raise ValueError("only head_prior_scaling = [1,1,1,1] is built (light_head_rfcn_eval.py:222)")
        shape = proposals_roi.shape
        _, boxes = ops.head_decode(proposals_roi.reshape(-1, 4), pred_location.reshape(-1, 4).contiguous(), 0, 4, 0)
        return boxes.reshape(shape)
