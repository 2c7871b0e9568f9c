"""ResNet-v2 (pre-activation) builders -- the API of the reference's ``net/resnet_v2.py``.

Same names and argument order (``batch_norm_relu``, ``fixed_padding``, ``conv2d_fixed_padding``,
``bottleneck_block``, ``block_layer``, ``imagenet_resnet_v2_generator``, ``imagenet_resnet_v2``), with one extra
trailing argument: the ``VariableStore`` that plays the role of TF's variable scope.  Tensors are NHWC bf16
CUDA tensors (the reference's ``data_format='channels_last'``); inference only here (``is_training=False``:
batch-norm uses its moving statistics, net/resnet_v2.py:41-50).

Compute: every convolution is ``ops.conv2d_nhwc`` (tcgen05 implicit GEMM); batch-norm + ReLU are folded into
the epilogue of the convolution that PRODUCES the tensor wherever the graph allows it:
  conv1x1 -> BN -> ReLU and conv3x3 -> BN -> ReLU inside a block are one kernel each;
  ``conv1x1 + shortcut`` writes the sum AND the next block's pre-activation ``relu(bn(sum))`` (second output).
Strided convolutions (``conv2d_fixed_padding`` with strides > 1, :89-100) are implicit GEMMs as well: the
tensor map reads every 2nd pixel and its out-of-bounds zero fill is ``fixed_padding`` (:62-86); the 3-channel
7x7/s2 stem runs in the kernel's fold_w mode on a row-padded NHWC8 copy of the image.

``lighthead_resnet50_body`` is the Light-Head composition SURVEY 8(a3) defines from these builders plus the
reference's own dilation pattern (net/xdet_body.py:28-121): block_layer1-3 as-is (stride 16) ->
batch_norm_relu -> RPN feature [N,h,w,1024]; block_layer4 with stride 1 and dilation 2 -> batch_norm_relu ->
[N,h,w,2048] for ``large_sep_kernel``.
"""
import torch

from .. import ops

_BATCH_NORM_DECAY = 0.997
_BATCH_NORM_EPSILON = 1e-5


def _bn(store, channels):
    return store.batch_norm(store.auto_name("batch_normalization"), channels)


def batch_norm_relu(inputs, is_training, data_format, store=None, bn=None):
    """Performs a batch normalization followed by a ReLU (net/resnet_v2.py:41-50)."""
    assert not is_training, "training-mode batch norm is not part of this build (see DESIGN.md)"
    assert data_format == "channels_last"
    bn = bn or _bn(store, inputs.shape[-1])
    scale, bias = store.folded_bn(bn, _BATCH_NORM_EPSILON)
    return ops.affine_relu(inputs, scale, bias, relu=True)


def fixed_padding(inputs, kernel_size, data_format):
    """Pads the input along the spatial dimensions independently of input size (net/resnet_v2.py:62-86)."""
    assert data_format == "channels_last"
    pad_total = kernel_size - 1
    pad_beg = pad_total // 2
    pad_end = pad_total - pad_beg
    return torch.nn.functional.pad(inputs, (0, 0, pad_beg, pad_end, pad_beg, pad_end))


def _conv_kernel(store, cin, filters, kernel_size, init=None):
    name = store.auto_name("conv2d")
    with store.scope(name):
        return store.get("kernel", (kernel_size, kernel_size, cin, filters), init or store.variance_scaling)


def _packed(store, kern, mode="conv"):
    """bf16 GEMM weight of a TF conv kernel [KH,KW,Cin,Cout]: 'conv' = tap-major with per-tap channel padding
    (stride-1 implicit GEMM), 'patch' = one K axis of KH*KW*Cin (for im2col'ed strided convs)."""
    key = ("w", kern[0], mode)
    if key not in store.derived:
        w = kern[1]
        kh, kw, cin, cout = w.shape
        if mode == "conv":
            store.derived[key] = ops.pack_conv_weight(w.permute(3, 2, 0, 1))
        else:
            store.derived[key] = ops.pack_conv_weight(w.reshape(kh * kw * cin, cout).t().reshape(cout, kh * kw * cin, 1, 1))
    return store.derived[key]


def _run_conv(store, x, kern, strides, dilation=1, **epilogue):
    """conv2d_fixed_padding semantics on NHWC bf16: SAME for stride 1, explicit pad + VALID otherwise
    (net/resnet_v2.py:89-100).  Strided convolutions read every 2nd pixel through the tensor map (no im2col)."""
    kh, kw, cin, cout = kern[1].shape
    if strides == 1:
        return ops.conv2d_nhwc(x, _packed(store, kern), cout, kh, kw, dilation=(dilation, dilation), padding="SAME",
                               cin=cin, **epilogue)
    N, H, W, _ = x.shape
    pad = (kh - 1) // 2  # fixed_padding: pad_beg = (k-1)//2, then VALID
    Ho = (H + (kh - 1) - kh) // strides + 1
    Wo = (W + (kw - 1) - kw) // strides + 1
    return ops.conv2d_nhwc(x, _packed(store, kern), cout, kh, kw, padding=(pad, pad, Ho, Wo),
                           strides=(strides, strides), cin=cin, **epilogue)


def conv2d_fixed_padding(inputs, filters, kernel_size, strides, data_format, kernel_initializer=None, store=None,
                         **epilogue):
    """Strided 2-D convolution with explicit padding (net/resnet_v2.py:89-100)."""
    assert data_format == "channels_last"
    kern = _conv_kernel(store, inputs.shape[-1], filters, kernel_size, kernel_initializer)
    return _run_conv(store, inputs, kern, strides, **epilogue)


def bottleneck_block(inputs, filters, is_training, projection_shortcut, strides, data_format, store=None,
                     dilation_rate=1, preact=None, next_bn=None, sum_unused=False, next_forms="pair"):
    """Bottleneck block variant for residual networks with BN before convolutions (net/resnet_v2.py:142-184;
    with ``dilation_rate`` > 1 the 3x3 is the dilated SAME conv of xdet_bottleneck_block, net/xdet_body.py:39-81).

    ``preact``: relu(bn(inputs)) if the producer of ``inputs`` already computed it; ``next_bn``: batch-norm of the
    consumer, to be fused as a second output; ``sum_unused``: nothing reads the raw sum (the consumer only takes
    relu(bn(sum))), so it is not stored.  Returns (sum_or_None, fused_next_preact_or_None).
    """
    assert not is_training and data_format == "channels_last"
    shortcut = inputs
    bn1 = _bn(store, preact.shape[-1] if inputs is None else inputs.shape[-1])
    if preact is None:
        preact = batch_norm_relu(inputs, is_training, data_format, store, bn=bn1)
    if projection_shortcut is not None:
        shortcut = projection_shortcut(preact)
    k1 = _conv_kernel(store, preact.shape[-1], filters, 1)
    bn2 = _bn(store, filters)
    s2, b2 = store.folded_bn(bn2, _BATCH_NORM_EPSILON)
    # ("f16x2" precision: the two inner activations are read by convolutions only -> stored as split planes only; the
    # sum is read as a residual only -> fp32 only)
    t = _run_conv(store, preact, k1, 1, scale=s2, bias=b2, relu=True, forms="pair")
    k2 = _conv_kernel(store, filters, filters, 3)
    bn3 = _bn(store, filters)
    s3, b3 = store.folded_bn(bn3, _BATCH_NORM_EPSILON)
    t = _run_conv(store, t, k2, strides if dilation_rate == 1 else 1, dilation=dilation_rate, scale=s3, bias=b3,
                  relu=True, forms="pair")
    k3 = _conv_kernel(store, filters, 4 * filters, 1)
    out2 = None
    ep = {"residual": shortcut, "forms": "f32"}
    if next_bn is not None:
        sn, bnb = store.folded_bn(next_bn, _BATCH_NORM_EPSILON)
        if ops.conv.PRECISION == "f16x2":  # the kernel wrapper allocates the forms that are needed
            ep.update(out2=True, scale2=sn, bias2=bnb, skip_out=bool(sum_unused), forms2=next_forms)
            return _run_conv(store, t, k3, 1, **ep)
        out2 = torch.empty_like(shortcut)
        ep.update(out2=out2, scale2=sn, bias2=bnb, skip_out=bool(sum_unused))
    y = _run_conv(store, t, k3, 1, **ep)
    return y, out2


def block_layer(inputs, filters, block_fn, blocks, strides, is_training, name, data_format, store=None,
                dilation_rate=1, preact=None, fuse_next=False, sum_unused=False, fused_forms="pair"):
    """Creates one layer of blocks for the ResNet model (net/resnet_v2.py:187-223; dilated form
    net/xdet_body.py:84-121).  Returns the layer output (sum of the last block); with ``fuse_next`` the
    batch-norm that follows this layer in creation order (the next layer's first pre-activation, or a trailing
    batch_norm_relu) is fused into the last convolution and (sum, relu(bn(sum))) is returned; ``sum_unused``
    then drops the raw sum (None is returned in its place).  ``fused_forms`` ("f16x2" precision only): the forms of
    that fused output -- "pair" when only convolutions read it, "both" when it is also a result of the model."""
    filters_out = 4 * filters

    def projection_shortcut(x):
        kern = _conv_kernel(store, x.shape[-1], filters_out, 1)
        return _run_conv(store, x, kern, strides if dilation_rate == 1 else 1, forms="f32")

    x, pre = inputs, preact
    for i in range(blocks):
        last = i + 1 == blocks
        # the consumer's first batch-norm is created by the consumer itself; peek its name to fuse it
        x, pre = block_fn(x, filters, is_training, projection_shortcut if i == 0 else None,
                          strides if i == 0 else 1, data_format, store=store, dilation_rate=dilation_rate, preact=pre,
                          next_bn=_peek_next_bn(store, filters_out) if (not last or fuse_next) else None,
                          sum_unused=last and fuse_next and sum_unused,
                          next_forms=fused_forms if last else "pair")
    return (x, pre) if fuse_next else x


def _peek_next_bn(store, channels, ahead=3):
    """Batch-norm variables the NEXT block will create first, created now (same automatic name) so that they can
    be fused into the convolution that produces its input.  A bottleneck block creates exactly three
    batch-norms, so seen from before a block the next block's first one is ``ahead`` = 3 names further."""
    c = store._counters[-1]
    i = c.get("batch_normalization", 0) + ahead
    name = "batch_normalization" if i == 0 else "batch_normalization_%d" % i
    return store.batch_norm(name, channels)  # does not advance the counter


def imagenet_resnet_v2_generator(block_fn, layers, num_classes, data_format=None, store=None):
    """Generator for ImageNet ResNet v2 models (net/resnet_v2.py:290-356), truncated after block_layer4 + the
    final batch_norm_relu: the classification head (avg-pool + dense) is not on the detector path."""
    data_format = data_format or "channels_last"

    def model(inputs, is_training):
        x = stem(inputs, store)
        x = block_layer(x, 64, block_fn, layers[0], 1, is_training, "block_layer1", data_format, store)
        x = block_layer(x, 128, block_fn, layers[1], 2, is_training, "block_layer2", data_format, store)
        x = block_layer(x, 256, block_fn, layers[2], 2, is_training, "block_layer3", data_format, store)
        x = block_layer(x, 512, block_fn, layers[3], 2, is_training, "block_layer4", data_format, store)
        return batch_norm_relu(x, is_training, data_format, store)

    return model


def imagenet_resnet_v2(resnet_size, num_classes, data_format=None, store=None):
    """Returns the ResNet model for a given size (net/resnet_v2.py:359-375); bottleneck sizes only."""
    model_params = {50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3], 200: [3, 24, 36, 3]}
    if resnet_size not in model_params:
        raise ValueError("Not a valid resnet_size:", resnet_size)
    return imagenet_resnet_v2_generator(bottleneck_block, model_params[resnet_size], num_classes, data_format, store)


def stem(image_nchw_f32, store, fuse_next=False, pooled_unused=False):
    """7x7/s2 initial conv with fixed padding + 3x3/s2 SAME max-pool (net/resnet_v2.py:320-328) on the
    fp32 NCHW image the input pipeline delivers; returns NHWC bf16 (with ``fuse_next`` also relu(bn(.)) of the
    first block's batch-norm, computed by the pooling kernel)."""
    N, C, H, W = image_nchw_f32.shape
    kern = _conv_kernel(store, C, 64, 7)
    key = ("w", kern[0], "fold")
    if key not in store.derived:
        store.derived[key] = ops.pack_fold_weight(kern[1].permute(3, 2, 0, 1))
    y = ops.conv2d_image_fold(image_nchw_f32.contiguous(), store.derived[key], 64, 7, 7, 2, 3, forms="f32")
    if not fuse_next:
        return ops.maxpool3x3s2_same(y)
    s1, b1 = store.folded_bn(_peek_next_bn(store, 64, ahead=0), _BATCH_NORM_EPSILON)
    # ("f16x2" precision: the first block has a projection shortcut, so only relu(bn(pooled)) is read, by convolutions)
    return ops.maxpool3x3s2_same(y, s1, b1, forms="none" if pooled_unused else "both", forms2="pair")


def lighthead_resnet50_body(image_nchw_f32, is_training, store, layers=(3, 4, 6, 3), after_rpn_feat=None):
    """Light-Head R-CNN backbone on ResNet-50 v2 (composition).
    -> (rpn_feature [N,h,w,1024], backbone_feature [N,h,w,2048]) NHWC bf16, both after batch_norm_relu.
    Every batch_norm_relu between layers is produced by the convolution (or pooling) kernel that writes its
    input; raw sums nobody else reads are never stored.  ``after_rpn_feat(rpn_feature)`` is called as soon as
    the RPN feature exists (the model_fn launches the RPN head and forks the proposal stream there).
    ``is_training=True``: ``store`` = the trainer (see net/xception_body.py, "training mode")."""
    if is_training:
        from .xception_body import _trainer
        tr = _trainer(store, "lighthead_resnet50_body")
        if tr.xception:
            raise ValueError("this trainer was built with backbone='xception'")
        rpn_feat = tr.fwd_backbone_mid()
        if after_rpn_feat is not None:
            after_rpn_feat(rpn_feat)
        return rpn_feat, tr.fwd_backbone_exit()
    df = "channels_last"
    x, pre = stem(image_nchw_f32, store, fuse_next=True, pooled_unused=True)
    # layers 2 and 3 start with a projection shortcut that reads relu(bn(x)): the raw x of layers 1 and 2 is unused
    _, pre = block_layer(x, 64, bottleneck_block, layers[0], 1, is_training, "block_layer1", df, store, preact=pre,
                         fuse_next=True, sum_unused=True)
    _, pre = block_layer(None, 128, bottleneck_block, layers[1], 2, is_training, "block_layer2", df, store,
                         preact=pre, fuse_next=True, sum_unused=True)
    # after layer 3 the next batch-norm in creation order is the RPN feature's batch_norm_relu
    x, rpn_feat = block_layer(None, 256, bottleneck_block, layers[2], 2, is_training, "block_layer3", df, store,
                              preact=pre, fuse_next=True, fused_forms="both")
    store.auto_name("batch_normalization")  # consumed by the fused second output above
    if after_rpn_feat is not None:
        after_rpn_feat(rpn_feat)
    _, backbone = block_layer(x, 512, bottleneck_block, layers[3], 1, is_training, "block_layer4", df, store,
                              dilation_rate=2, fuse_next=True, sum_unused=True, fused_forms="both")
    store.auto_name("batch_normalization")
    return rpn_feat, backbone
