"""``AnchorCreator`` / ``AnchorEncoder`` -- the slice of the reference's
``preprocessing/anchor_manipulator.py`` that the inference hot path uses, same constructor arguments.

* ``AnchorCreator.get_layer_anchors`` / ``get_all_anchors`` (reference :698-757): anchor centres and sizes as
  fp32 constants, computed once on the host exactly as the reference does (python-double sqrt, then cast).
* ``AnchorEncoder.decode_all_anchors`` (:641-669) and ``ext_decode_rois`` (:671-683): CUDA kernels
  ``xdet_rpn_decode`` / ``xdet_head_decode``.
* ``AnchorEncoder.encode_all_anchors`` (:118-171, :319-335) and ``ext_encode_rois`` (:337-432), the training
  targets: the reference's method names over the kernels the training step launches directly
  (``xdet_match_encode``, ``xdet_sample_fg_bg``; ``tf.random_shuffle`` = stable argsort of injected keys).
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

    # ---- training targets -----------------------------------------------------------------------------------
    def _flat_anchors(self, index=0):
        """All anchors of one layer, (y, x, a) order with a fastest: point form [A_tot,4] (center2point) and centre
        form [A_tot,4]."""
        yref, xref, href, wref = self._dev_anchors[index]
        cells, A = yref.numel(), href.numel()
        cy = yref.reshape(cells, 1).expand(cells, A).reshape(-1)
        cx = xref.reshape(cells, 1).expand(cells, A).reshape(-1)
        hh = href.reshape(1, A).expand(cells, A).reshape(-1)
        ww = wref.reshape(1, A).expand(cells, A).reshape(-1)
        pt = torch.stack([cy - hh / 2., cx - ww / 2., cy + hh / 2., cx + ww / 2.], -1).contiguous()
        return pt, torch.stack([cy, cx, hh, ww], -1).contiguous()

    def encode_all_anchors(self, labels, bboxes):
        """Reference :319-335 (``encode_anchor`` :118-171 per layer): labels [G] / [N,G] (<= 0: padding), bboxes
        [G,4] / [N,G,4] CUDA tensors -> (ground_labels, anchor_regress_targets, ground_scores, ground_bboxes,
        n_layers), lists with one entry per feature layer: labels int32 (>0 matched class, 0 background, -1 ignored),
        targets fp32 [...,4], scores fp32, each over the layer's anchors ([A_tot] or [N,A_tot]); ground_bboxes = the
        anchors in point form."""
        from ..ops import train as T
        single = labels.dim() == 1
        gl = (labels.reshape(1, -1) if single else labels).to(torch.int32).contiguous()
        gb = (bboxes.reshape(1, -1, 4) if single else bboxes).to(torch.float32).contiguous()
        ground_labels, targets, scores, points = [], [], [], []
        for index in range(len(self._anchors)):
            pt, yxhw = self._flat_anchors(index)
            l, t, s = T.match_encode(pt, gb, gl, float(self._allowed_borders[index]), self._positive_threshold,
                                     self._ignore_threshold, ref_yxhw=yxhw)
            ground_labels.append(l[0] if single else l)
            targets.append(t[0] if single else t)
            scores.append(s[0] if single else s)
            points.append(pt)
        return ground_labels, targets, scores, points, len(self._anchors)

    def ext_encode_rois(self, all_rois, all_labels, all_bboxes, rois_per_image, fg_fraction, allowed_border,
                        head_prior_scaling=[1., 1., 1., 1.], keys=None):
        """Reference :337-432: append the ground-truth boxes to the RoIs, match (fg >= rpn_fg_thres, bg below
        rpn_bg_high_thres and above rpn_bg_low_thres), sample ``rois_per_image`` per image at ``fg_fraction``
        foreground.  all_rois [N,R,4], all_labels [N,G] (<= 0: padding), all_bboxes [N,G,4] -> (rois [N,S,4],
        targets [N,S,4], labels [N,S] int32, scores [N,S]).  ``keys``: {'roi_fg','roi_bg': [N,R+G], 'roi_up': [N,S]}
        uniform fp32 arrays standing in for the three tf.random_shuffle calls (drawn here when omitted)."""
        from ..ops import train as T
        if any(float(p) != 1.0 for p in head_prior_scaling):
            raise ValueError("only head_prior_scaling = [1,1,1,1] is built (light_head_rfcn_train.py:252)")
        N, S = all_rois.shape[0], int(rois_per_image)
        gl = all_labels.to(torch.int32).contiguous()
        gb = all_bboxes.to(torch.float32).contiguous()
        rois_all = torch.cat([all_rois, gb * (gl > 0).unsqueeze(-1).float()], dim=1).contiguous()
        n = rois_all.shape[1]
        if keys is None:
            keys = {'roi_fg': torch.rand((N, n), device=rois_all.device), 'roi_bg': torch.rand((N, n), device=rois_all.device),
                    'roi_up': torch.rand((N, S), device=rois_all.device)}
        lab, tgt, sc = T.match_encode(rois_all, gb, gl, float(allowed_border), self._rpn_fg_thres,
                                      self._rpn_bg_high_thres)
        # the reference appends only the valid ground-truth boxes (tf.boolean_mask, :345-347); the padded slots ride
        # along here as zero boxes and are marked 'ignore' so that no threshold setting can sample them
        lab[:, lab.shape[1] - gl.shape[1]:].masked_fill_(gl <= 0, -1)
        idx, _ = T.sample_fg_bg(lab, sc, float(self._rpn_bg_low_thres), int(round(S * fg_fraction)), S,
                                keys['roi_fg'], keys['roi_bg'], keys['roi_up'])
        idx = idx.long()
        gather4 = idx.unsqueeze(-1).expand(N, S, 4)
        return (torch.gather(rois_all, 1, gather4).contiguous(), torch.gather(tgt, 1, gather4).contiguous(),
                torch.gather(lab, 1, idx).contiguous(), torch.gather(sc, 1, idx).contiguous())
