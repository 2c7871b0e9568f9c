"""Light-Head R-CNN head builders -- the API of the reference's ``net/xception_body.py``
(``get_rpn`` :381-400, ``get_proposals`` :402-448, ``large_sep_kernel`` :450-475, ``get_head`` :477-560),
same names and argument order plus a trailing ``VariableStore``.  NHWC bf16 tensors in, inference only.

Fusions relative to the reference graph (all exact re-associations of the same linear algebra):
  * get_rpn: the two 1x1 heads (2A class logits, 4A box deltas) are ONE GEMM with 6A output channels;
  * large_sep_kernel: the two branches read the same input, so the two 15x1 convs are one conv to 2*mid
    channels and ``branch_0b + branch_1b`` is one 1x15 conv over the concatenated 2*mid channels (K = 15*2*mid);
    the trailing batch_norm_relu is folded into its epilogue, which writes fp32 NCHW -- the layout and dtype
    PsRoIAlign's contract requires;
  * get_head: fc_cls and fc_loc are one GEMM with num_classes + 4 outputs.
``XceptionBody`` (:236-379) runs every separable convolution as a depthwise bandwidth kernel (its leading ReLU
applied while loading) + a pointwise tcgen05 GEMM whose epilogue carries the batch-norm, the block's residual
add and, where the graph wants it, a ReLU'd second output; the entry flow's ``max_pool + residual`` is one kernel.
"""
import torch

from .. import ops
from . import resnet_v2

USE_FUSED_BN = True
BN_EPSILON = 0.0001
BN_MOMENTUM = 0.99


# ---- training mode -----------------------------------------------------------------------------------------------
# With is_training=True the builders below do what they do in the reference's training graph
# (light_head_rfcn_train.py:289-407): they run their piece of the step in training mode (batch statistics, tape for the
# explicit backward).  The object that owns the variables, the tape and the gradients is the trainer
# (light_head_rfcn_train.LightHeadTrainer): pass it as ``store``, open the step with ``trainer.begin_step(images,
# gt_boxes, gt_labels, keys)`` (the Estimator feeding features / labels), call the builders in the reference's order, and
# finish with ``trainer.backward(); trainer.apply_gradients()`` (optimizer.minimize).  ``LightHeadTrainer.step`` is the
# same pieces with the proposal work forked onto a second stream.
def _trainer(store, what):
    if store is None or not hasattr(store, "fwd_backbone_mid"):
        raise TypeError("%s(is_training=True) needs the trainer (light_head_rfcn_train.LightHeadTrainer) as `store`: it "
                        "owns the variables, the forward tape and the gradient buffer of the training graph" % what)
    if getattr(store, "t", None) is None:
        raise RuntimeError("%s(is_training=True): call trainer.begin_step(images, gt_boxes, gt_labels, keys) first" % what)
    return store


def _as_format(x_nhwc, data_format):
    """Builders hand activations on as NHWC; 'channels_first' callers get the NCHW view of the same memory."""
    if data_format == "channels_first":
        return x_nhwc.permute(0, 3, 1, 2)
    assert data_format == "channels_last", data_format
    return x_nhwc


def _conv_vars(store, cin, filters, kh, kw, init=None, use_bias=True):
    name = store.auto_name("conv2d")
    with store.scope(name):
        k = store.get("kernel", (kh, kw, cin, filters), init or store.glorot_normal)
        b = store.get("bias", (filters,), store.zeros) if use_bias else None
    return k, b


def _dense_vars(store, name, cin, units):
    with store.scope(name):
        k = store.get("kernel", (cin, units), store.glorot_normal)
        b = store.get("bias", (units,), store.zeros)
    return k, b


def _derived(store, key, fn):
    if key not in store.derived:
        store.derived[key] = fn()
    return store.derived[key]


def _bn_named(store, name, channels):
    return store.folded_bn(store.batch_norm(name, channels), BN_EPSILON)


def _conv_named(store, name, cin, filters, k):
    with store.scope(name):
        return store.get("kernel", (k, k, cin, filters), store.glorot_normal)


def _sep_vars(store, name, cin, filters):
    with store.scope(name):
        dw = store.get("depthwise_kernel", (3, 3, cin, 1), store.glorot_normal)
        pw = store.get("pointwise_kernel", (1, 1, cin, filters), store.glorot_normal)
    return dw, pw


def _separable(store, x, name, filters, relu_in, dilation=1, bn_name=None, **epilogue):
    """tf.layers.separable_conv2d(3x3, SAME, use_bias=False) + batch_normalization ``name + '_bn'``
    (reference :220-234): depthwise kernel (ReLU on load when ``relu_in``) -> pointwise GEMM with the BN folded
    into the epilogue (plus whatever ``epilogue`` adds: relu / residual / out2 ...)."""
    cin = x.shape[-1]
    dw, pw = _sep_vars(store, name, cin, filters)
    w9 = _derived(store, ("dw", dw[0]), lambda: dw[1].reshape(9, cin).float().contiguous())
    wp = _derived(store, ("w", pw[0]), lambda: ops.pack_conv_weight(pw[1].permute(3, 2, 0, 1)))
    scale, bias = _bn_named(store, bn_name or (name + "_bn"), filters)
    t = ops.depthwise3x3(x, w9, dilation=dilation, relu_in=relu_in)
    epilogue.setdefault("forms", "f32")  # ("f16x2" precision) the usual reader is the next depthwise kernel
    return ops.conv2d_nhwc(t, wp, filters, 1, 1, scale=scale, bias=bias, cin=cin, **epilogue)


def relu_separable_bn_block(inputs, filters, name_prefix, is_training, data_format, store=None, **epilogue):
    """relu -> separable_conv2d -> batch_normalization (reference :220-234)."""
    assert data_format == "channels_last" and not is_training
    return _separable(store, inputs, name_prefix, filters, relu_in=True, **epilogue)


def XceptionBody(input_image, num_classes, is_training=False, data_format='channels_last', store=None,
                 after_mid=None):
    """Xception backbone at stride 16 with the last two separable convs atrous (reference :236-379).
    ``input_image``: fp32 NCHW [N,3,H,W] (what the input pipeline delivers).  Returns
    (mid_outputs [N,h,w,728], outputs [N,h,w,2048]) NHWC bf16, both ReLU'd as in the reference.
    ``after_mid(mid_outputs)`` is called as soon as the RPN feature exists (the model_fn forks there).
    ``is_training=True`` (``store`` = the trainer): batch statistics, tape; both data formats."""
    if is_training:
        tr = _trainer(store, "XceptionBody")
        if not tr.xception:
            raise ValueError("this trainer was built with backbone=%r" % tr.params['backbone'])
        assert input_image is tr.t.images or input_image.data_ptr() == tr.t.images.data_ptr()
        mid = tr.fwd_backbone_mid()
        if after_mid is not None:
            after_mid(mid)
        return _as_format(mid, data_format), _as_format(tr.fwd_backbone_exit(), data_format)
    assert data_format == "channels_last"
    df = data_format
    # ---- entry flow: two VALID 3x3 convs (the first strided, on the 3-channel image: fold_w mode) ----
    k = _conv_named(store, "block1_conv1", input_image.shape[1], 32, 3)
    wf = _derived(store, ("w", k[0], "fold"), lambda: ops.pack_fold_weight(k[1].permute(3, 2, 0, 1)))
    sc, bi = _bn_named(store, "block1_conv1_bn", 32)
    x = ops.conv2d_image_fold(input_image.contiguous(), wf, 32, 3, 3, 2, 0, scale=sc, bias=bi, relu=True, forms="pair")
    k = _conv_named(store, "block1_conv2", 32, 64, 3)
    w = _derived(store, ("w", k[0]), lambda: ops.pack_conv_weight(k[1].permute(3, 2, 0, 1)))
    sc, bi = _bn_named(store, "block1_conv2_bn", 64)
    x = ops.conv2d_nhwc(x, w, 64, 3, 3, padding="VALID", scale=sc, bias=bi, relu=True)

    def strided_residual(x, idx, filters):
        """1x1 / stride 2 'same' conv + batch-norm on the block input (:262-266, :291-296, :311-316)."""
        k = _conv_named(store, "conv2d_%d" % idx, x.shape[-1], filters, 1)
        w = _derived(store, ("w", k[0]), lambda: ops.pack_conv_weight(k[1].permute(3, 2, 0, 1)))
        sc, bi = _bn_named(store, "batch_normalization_%d" % idx, filters)
        N, H, W, _ = x.shape
        return ops.conv2d_nhwc(x, w, filters, 1, 1, padding=(0, 0, -(-H // 2), -(-W // 2)), strides=(2, 2), scale=sc,
                               bias=bi, forms="f32")

    residual = strided_residual(x, 1, 128)
    x = _separable(store, x, "block2_sepconv1", 128, relu_in=False)  # its input is already ReLU'd (:268)
    x = relu_separable_bn_block(x, 128, "block2_sepconv2", is_training, df, store)
    x = ops.maxpool3x3s2_same(x, residual=residual)  # block2_pool + residual_add_0
    for blk, idx, filters in ((3, 2, 256), (4, 3, 728)):
        residual = strided_residual(x, idx, filters)
        x = relu_separable_bn_block(x, filters, "block%d_sepconv1" % blk, is_training, df, store)
        x = relu_separable_bn_block(x, filters, "block%d_sepconv2" % blk, is_training, df, store)
        x = ops.maxpool3x3s2_same(x, residual=residual)
    # ---- middle flow: 8 x (3 x relu-sepconv-bn) + identity (:325-334) ----
    mid_outputs = None
    for index in range(8):
        prefix = "block%d" % (index + 5)
        residual = x
        t = relu_separable_bn_block(x, 728, prefix + "_sepconv1", is_training, df, store)
        t = relu_separable_bn_block(t, 728, prefix + "_sepconv2", is_training, df, store)
        if index < 7:
            x = relu_separable_bn_block(t, 728, prefix + "_sepconv3", is_training, df, store, residual=residual)
        else:  # the last sum is also needed ReLU'd: mid_outputs = relu(x) (before_block13_act, :336)
            mid_outputs = torch.empty_like(residual)
            one, zero = _derived(store, ("unit", 728), lambda: (torch.ones(728, device=x.device),
                                                                  torch.zeros(728, device=x.device)))
            x = relu_separable_bn_block(t, 728, prefix + "_sepconv3", is_training, df, store, residual=residual,
                                        out2=mid_outputs, scale2=one, bias2=zero, forms="both")
    if after_mid is not None:
        after_mid(mid_outputs)
    # ---- exit flow with the stride removed and dilation 2 in block14 (:337-376) ----
    k = _conv_named(store, "conv2d_4", 728, 1024, 1)
    w = _derived(store, ("w", k[0]), lambda: ops.pack_conv_weight(k[1].permute(3, 2, 0, 1)))
    sc, bi = _bn_named(store, "batch_normalization_4", 1024)
    residual = ops.conv2d_nhwc(x, w, 1024, 1, 1, scale=sc, bias=bi, forms="f32")
    t = relu_separable_bn_block(x, 728, "block13_sepconv1", is_training, df, store)
    x = relu_separable_bn_block(t, 1024, "block13_sepconv2", is_training, df, store, residual=residual)
    x = _separable(store, x, "block14_sepconv1", 1536, relu_in=False, dilation=2, relu=True)
    outputs = _separable(store, x, "block14_sepconv2", 2048, relu_in=False, dilation=2, relu=True, forms="both")
    return mid_outputs, outputs


def get_rpn(net_input, num_anchors, is_training, data_format, var_scope, store=None):
    """3x3 SAME conv -> 512 + bias + ReLU, then two 1x1 convs -> 2A / 4A (+bias) (reference :381-400).
    Returns ONE fp32 NHWC tensor [N,h,w,6A]: class logits in channels [0,2A), box deltas in [2A,6A)
    (``rpn_cls_score, rpn_bbox_pred`` are its two channel slices).  ``is_training=True``: ``store`` = the trainer."""
    if is_training:
        tr = _trainer(store, "get_rpn")
        return tr.fwd_rpn(tr.t.rpn_feat)
    assert data_format == "channels_last"
    cin = net_input.shape[-1]
    with store.scope(var_scope):
        k0, b0 = _conv_vars(store, cin, 512, 3, 3)
        k1, b1 = _conv_vars(store, 512, 2 * num_anchors, 1, 1)
        k2, b2 = _conv_vars(store, 512, 4 * num_anchors, 1, 1)
    w0 = _derived(store, ("w", k0[0]), lambda: ops.pack_conv_weight(k0[1].permute(3, 2, 0, 1)))
    rpn_relu = ops.conv2d_nhwc(net_input, w0, 512, 3, 3, bias=b0[1], relu=True, forms="pair")
    w12 = _derived(store, ("w", k1[0], k2[0]),
                   lambda: ops.pack_conv_weight(torch.cat([k1[1], k2[1]], dim=3).permute(3, 2, 0, 1)))
    b12 = _derived(store, ("b", b1[0], b2[0]), lambda: torch.cat([b1[1], b2[1]]).contiguous())
    return ops.conv2d_nhwc(rpn_relu, w12, 6 * num_anchors, 1, 1, bias=b12, out_layout="nhwc_f32", forms="f32")


def get_proposals(object_score, bboxes_pred, encode_fn, rpn_pre_nms_top_n, rpn_post_nms_top_n, nms_threshold,
                  rpn_min_size, is_training, data_format, shuffle_keys=None, store=None):
    """clip -> filter/top-k -> NMS -> upsample (reference :402-448), on the GPU instead of /cpu:0.
    object_score [N,A], bboxes_pred [N,A,4].  Inference: returns the proposal boxes [N,post,4].
    ``is_training=True`` (``store`` = the trainer): also the ``encode_fn`` step (ext_encode_rois with the trainer's
    thresholds and shuffle keys) -> (proposals_bboxes, proposals_targets, proposals_labels, proposals_scores) (:445-448)."""
    if is_training:
        tr = _trainer(store, "get_proposals")
        return tr.fwd_proposals_and_targets()  # (runs the model_fn's objectness / decode first if it has not happened)
    rois, _, _ = ops.rpn_select(object_score, bboxes_pred, rpn_pre_nms_top_n, rpn_post_nms_top_n, nms_threshold,
                                rpn_min_size, shuffle_keys)
    return rois


def large_sep_kernel(net_input, depth_mid, depth_output, is_training, data_format, var_scope, store=None):
    """Two branches of (15x1 conv -> depth_mid, 1x15 conv -> depth_output), summed, batch_norm_relu
    (reference :450-475).  Returns the thin feature map as fp32 NCHW [N,depth_output,h,w].
    ``is_training=True``: ``store`` = the trainer (batch statistics for the trailing batch_norm_relu)."""
    if is_training:
        tr = _trainer(store, "large_sep_kernel")
        assert depth_mid == 256 and depth_output == 490, "the trainer is built for the reference's 256 / 10*7*7"
        return tr.fwd_thin(tr.t.backbone)
    assert data_format == "channels_last"
    cin = net_input.shape[-1]
    with store.scope(var_scope):
        with store.scope("Branch_0"):
            a0k, a0b = _conv_vars(store, cin, depth_mid, 15, 1)
            b0k, b0b = _conv_vars(store, depth_mid, depth_output, 1, 15)
        with store.scope("Branch_1"):
            a1k, a1b = _conv_vars(store, cin, depth_mid, 15, 1)
            b1k, b1b = _conv_vars(store, depth_mid, depth_output, 1, 15)
        bn = store.batch_norm(store.auto_name("batch_normalization"), depth_output)
    wa = _derived(store, ("w", a0k[0], a1k[0]),
                  lambda: ops.pack_conv_weight(torch.cat([a0k[1], a1k[1]], dim=3).permute(3, 2, 0, 1)))
    ba = _derived(store, ("b", a0b[0], a1b[0]), lambda: torch.cat([a0b[1], a1b[1]]).contiguous())
    mid = ops.conv2d_nhwc(net_input, wa, 2 * depth_mid, 15, 1, bias=ba, forms="pair")
    wb = _derived(store, ("w", b0k[0], b1k[0]),
                  lambda: ops.pack_conv_weight(torch.cat([b0k[1], b1k[1]], dim=2).permute(3, 2, 0, 1)))
    scale, shift = store.folded_bn(bn, resnet_v2._BATCH_NORM_EPSILON)
    bb = _derived(store, ("b", b0b[0], b1b[0], "bn"), lambda: ((b0b[1] + b1b[1]) * scale + shift).contiguous())
    return ops.conv2d_nhwc(mid, wb, depth_output, 1, 15, scale=scale, bias=bb, relu=True, out_layout="nchw_f32")


def _point2center(proposals_bboxes):
    ymin, xmin, ymax, xmax = (proposals_bboxes[:, :, 0], proposals_bboxes[:, :, 1], proposals_bboxes[:, :, 2],
                              proposals_bboxes[:, :, 3])
    height, width = (ymax - ymin), (xmax - xmin)
    return torch.stack([ymin + height / 2., xmin + width / 2., height, width], dim=-1)


def get_head(net_input, pooling_op, grid_width, grid_height, loss_func, proposals_bboxes, num_classes, is_training,
             using_ohem, ohem_roi_one_image, data_format, var_scope, store=None, yxhw_bboxes=None,
             return_fused=False):
    """PS-RoI pooling + fc 2048 (ReLU) + fc_cls / fc_loc (reference :477-560), inference branch.
    net_input: thin feature map, fp32 NCHW (PsRoiAlign's contract).  Returns (cls_score [N,R,num_classes],
    bboxes_reg [N,R,4]) fp32.  ``is_training=True`` (``store`` = the trainer): the training branch with OHEM
    (:504-533: no-grad pass, per-RoI loss, top-k, the axis-1 gather) -> (cls_score, bboxes_reg) of the selected rows; the
    head loss (``loss_func`` of the reference = head_loss_func, train:381-399) is ``trainer.t.head_loss``."""
    if is_training or using_ohem:
        tr = _trainer(store, "get_head")
        if bool(using_ohem) != bool(tr.params['using_ohem']) or (using_ohem and ohem_roi_one_image !=
                                                                    tr.params['ohem_roi_one_image']):
            raise ValueError("using_ohem / ohem_roi_one_image differ from the trainer's parameters")
        cls_score, bboxes_reg, _ = tr.fwd_head()
        return cls_score, bboxes_reg
    if yxhw_bboxes is None:
        yxhw_bboxes = _point2center(proposals_bboxes)  # fp32 elementwise, same op order as the reference
    psroipooled_rois, _ = pooling_op(net_input, yxhw_bboxes.contiguous(), grid_width, grid_height)
    N, R = psroipooled_rois.shape[:2]
    feat = psroipooled_rois.reshape(N * R, -1)  # tf.reshape(pooled_feat, [-1, 10*gw*gh]) (:499)
    cin = feat.shape[1]
    with store.scope(var_scope):
        k1, b1 = _dense_vars(store, "subnet_fc", cin, 2048)
        kc, bc = _dense_vars(store, "fc_cls", 2048, num_classes)
        kl, bl = _dense_vars(store, "fc_loc", 2048, 4)
    w1 = _derived(store, ("w", k1[0]), lambda: ops.pack_conv_weight(k1[1].t().reshape(2048, cin, 1, 1)))
    if ops.conv.PRECISION in ("f16x2", "fp32x3"):  # fp32 activations: the pooled features are split inside conv2d_nhwc
        h = ops.conv2d_nhwc(feat.reshape(1, 1, N * R, cin), w1, 2048, 1, 1, bias=b1[1], relu=True, forms="pair")
    else:
        pitch = (cin + 7) // 8 * 8  # TMA rows must be 16-byte multiples
        a = ops.f32_to_bf16_rows(feat, pitch)
        h = ops.conv2d_nhwc(a.reshape(1, 1, N * R, pitch), w1, 2048, 1, 1, bias=b1[1], relu=True, cin=cin)
    w2 = _derived(store, ("w", kc[0], kl[0]),
                  lambda: ops.pack_conv_weight(torch.cat([kc[1], kl[1]], dim=1).t().reshape(num_classes + 4, 2048, 1, 1)))
    b2 = _derived(store, ("b", bc[0], bl[0]), lambda: torch.cat([bc[1], bl[1]]).contiguous())
    out = ops.conv2d_nhwc(h, w2, num_classes + 4, 1, 1, bias=b2, out_layout="nhwc_f32",
                          forms="f32").reshape(N, R, num_classes + 4)
    if return_fused:  # also the [N,R,num_classes+4] tensor both results are views of
        return out[..., :num_classes], out[..., num_classes:], out
    return out[..., :num_classes], out[..., num_classes:]
