"""Training-mode ``XceptionBody`` (net/xception_body.py:220-379 with is_training=True) on the CUDA kernels: forward
tape + explicit backward, the device-side twin of the CPU blueprint ``oracle/xception_backward.py`` (which is held to
autograd for all 154 trainable variables).  Same classes, same order of operations; every method is one or two
kernel launches:

  Conv.bwd        xdet_conv2d_wgrad_bf16 + the forward kernel on dY with flipped weights (input gradient)
  Depthwise.bwd   xdet_depthwise3x3_wgrad_bf16 + the forward kernel on dY with flipped taps
                  (+ xdet_relu_bwd_bf16 when the block ReLUs its input)
  BatchNorm       xdet_col_stats / xdet_bn_finalize forward, xdet_bn_relu_bwd_bf16 backward (relu flag 0 / 1)
  MaxPool         xdet_maxpool3x3s2_argmax_bf16 / xdet_maxpool3x3s2_bwd_bf16

``XceptionBodyTraining`` returns gradients as a dict in TF variable layouts (compared tensor by tensor with the
blueprint in tests/test_xception_train_gpu.py); ``TrainableXceptionBody`` is the same backbone over
``LightHeadTrainer``'s parameter classes (flat all-reduce buffer, momentum slots, moving statistics).
"""
import torch

from .. import _native, ops
from ..ops import train as T

BN_EPSILON = 0.0001


def _st():
    return torch.cuda.current_stream().cuda_stream


def depthwise3x3_wgrad(x, dy, dilation, relu_in, out=None):
    """dW [9, C] fp32 of the depthwise 3x3 'same' convolution (csrc/depthwise_wgrad.cu).  ``out``: 9*C fp32 elements the
    kernel ADDS the gradient to (a view of the trainer's flat gradient buffer: no temporary, no separate add)."""
    lib = _native.lib()
    N, H, W, C = x.shape
    if out is not None:
        assert out.dtype == torch.float32 and out.numel() == 9 * C and out.is_contiguous()
        dw = out
    else:
        dw = torch.zeros((9, C), dtype=torch.float32, device=x.device)
    fn = lib.xdet_depthwise3x3_wgrad_f32 if x.dtype == torch.float32 else lib.xdet_depthwise3x3_wgrad_bf16
    assert x.dtype == dy.dtype and x.is_contiguous() and dy.is_contiguous()
    _native.check(fn(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, H, W, C, dilation, 1 if relu_in else 0, _st()))
    return dw


BN_MOMENTUM = 0.99


def _image_nhwc(images):
    """fp32 NCHW image -> the NHWC tensor the first convolution reads: bf16 rows of 8 channels, or (fp32-accurate mode)
    fp32 NHWC as it is."""
    if ops.conv.PRECISION == "f16x2":
        return images.permute(0, 2, 3, 1).contiguous()
    if images.is_cuda and images.shape[1] <= 8 and images.dtype == torch.float32:   # (the stem's coalesced repack)
        return ops.image_to_nhwc8(images.contiguous(), 0, images.shape[3])
    return T.nchw_f32_to_nhwc_bf16(images.contiguous(), pitch=8)


def bn_relu_bwd_into(dy, x, st, relu, grad_view):
    """xdet_bn_relu_bwd_bf16 with its column sums written straight into ``grad_view`` ([0,C) = dbeta, [C,2C) = dgamma:
    the (beta, gamma) order of the trainer's VecParam; the flat gradient buffer is cleared when the step begins and this is
    the view's only writer, so the kernel need not clear it again)."""
    return T.bn_relu_bwd_into(dy, x, st, relu, grad_view, sums_zeroed=True)


class Conv(object):
    """tf.layers.conv2d without bias on NHWC bf16; ``w`` = TF kernel [kh,kw,cin,cout] fp32."""

    def __init__(self, name, w, stride=1, padding="SAME", need_dgrad=True):
        self.name, self.stride, self.padding = name, stride, padding
        self.kh, self.kw, self.cin, self.cout = w.shape
        w4 = w.permute(3, 2, 0, 1).contiguous()
        self.pack = ops.pack_conv_weight(w4)
        self.dpack = ops.pack_dgrad_weight(w4) if need_dgrad else None

    def _geom(self, H, W):
        s = self.stride
        if self.padding == "VALID":
            return (0, 0, (H - self.kh) // s + 1, (W - self.kw) // s + 1)
        if s == 1:
            return "SAME"
        assert self.kh == 1 and self.kw == 1    # the strided 1x1 'same' projections of the entry flow
        return (0, 0, -(-H // s), -(-W // s))

    def fwd(self, x):
        self.x, self.in_hw = x, tuple(x.shape[1:3])
        self.geom = self._geom(*self.in_hw)
        # every convolution of the body feeds a batch-norm: under the trainer (bf16 precision) the kernel's epilogue
        # accumulates that layer's batch statistics (they ride on the output tensor as ``_bn_sums``)
        arena, sums = getattr(self, "arena", None), None
        if arena is not None and x.dtype == torch.bfloat16 and self.cout % 8 == 0 and ops.conv.PRECISION == "bf16":
            sums = arena.take(2 * self.cout)
        y = ops.conv2d_nhwc(x, self.pack, self.cout, self.kh, self.kw, padding=self.geom,
                            strides=(self.stride, self.stride), cin=self.cin, **({} if sums is None else {"stats": sums}))
        if sums is not None:
            y._bn_sums = sums
        return y

    def bwd(self, dy, grads, leaf="kernel"):
        p = getattr(self, "p", None)     # TrainableXceptionBody: accumulate into the trainer's flat gradient buffer
        if p is not None:
            dw = p.dw
        else:
            cin_pad = (self.cin + 63) // 64 * 64
            dw = torch.zeros((self.cout, self.kh * self.kw, cin_pad), dtype=torch.float32, device=dy.device)
        wg = getattr(self, "wg_stream", None)   # the trainer's weight-gradient stream (joined by LightHeadTrainer.backward)
        if wg is not None and p is not None and self.dpack is not None and dy.dtype == torch.bfloat16:
            wg.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(wg):
                ops.conv2d_wgrad(self.x, dy, self.kh, self.kw, padding=self.geom, strides=(self.stride, self.stride),
                                 cin=self.cin, cout=self.cout, dw=dw)
            dy.record_stream(wg)
            wg.xdet_forked = True
        else:
            ops.conv2d_wgrad(self.x, dy, self.kh, self.kw, padding=self.geom, strides=(self.stride, self.stride),
                             cin=self.cin, cout=self.cout, dw=dw)
        if p is None:
            grads[self.name + "/" + leaf] = dw.view(self.cout, self.kh, self.kw, -1)[..., :self.cin].permute(
                1, 2, 3, 0).contiguous()
        if self.dpack is None:
            return None
        return ops.conv2d_dgrad(dy, self.dpack, self.cin, self.kh, self.kw, self.in_hw, padding=self.geom,
                                strides=(self.stride, self.stride), cout=self.cout)


class Depthwise(object):
    """Depthwise stage of tf.layers.separable_conv2d; ``w`` = [3,3,C,1] fp32."""

    def __init__(self, name, w, dil=1, relu_in=True):
        self.name, self.dil, self.relu_in = name, dil, relu_in
        self.C = w.shape[2]
        self.w9 = w.reshape(9, self.C).float().contiguous()

    def fwd(self, x):
        self.x = x
        return ops.depthwise3x3(x, self.w9, dilation=self.dil, relu_in=self.relu_in)

    def bwd(self, dy, grads):
        dy = dy.contiguous()
        vec = getattr(self, "vec", None)
        if vec is not None:
            depthwise3x3_wgrad(self.x, dy, self.dil, self.relu_in, out=vec.grad[:9 * self.C])
        else:
            dw = depthwise3x3_wgrad(self.x, dy, self.dil, self.relu_in)
            grads[self.name + "/depthwise_kernel"] = dw.reshape(3, 3, self.C, 1)
        # tap (kh,kw) -> (2-kh,2-kw); flipped per call: w9 may be a view of a master the optimizer updates
        da = ops.depthwise3x3(dy, self.w9.flip(0).contiguous(), dilation=self.dil, relu_in=False, forms="f32")
        return T.relu_bwd(da, self.x) if self.relu_in else da


class BatchNorm(object):
    """tf.layers.batch_normalization(training=True) (+ ReLU when ``relu``); moving statistics are not updated here."""

    def __init__(self, name, gamma, beta, relu=False):
        self.name, self.gamma, self.beta, self.relu = name, gamma, beta, relu

    def fwd(self, x):
        self.x = x
        moving = getattr(self, "moving", (None, None))   # TrainableXceptionBody: also update the moving statistics
        sums = getattr(x, "_bn_sums", None)
        if sums is not None and x.dtype == torch.bfloat16 and x.shape[-1] == self.gamma.numel():
            y, self.st = T.bn_train_apply(x, sums, self.gamma, self.beta, BN_EPSILON,
                                          None if moving[0] is None else BN_MOMENTUM, *moving, relu=self.relu)
            return y
        self.st = T.bn_train(x, self.gamma, self.beta, BN_EPSILON, None if moving[0] is None else BN_MOMENTUM, *moving)
        return ops.affine_relu(x, self.st.scale, self.st.shift, relu=self.relu)

    def bwd(self, dy, grads):
        vec = getattr(self, "vec", None)
        if vec is not None:
            return bn_relu_bwd_into(dy.contiguous(), self.x, self.st, self.relu, vec.grad)
        dx, dgamma, dbeta = T.bn_relu_bwd(dy.contiguous(), self.x, self.st, relu=self.relu)
        grads[self.name + "/gamma"], grads[self.name + "/beta"] = dgamma, dbeta
        return dx


class MaxPool(object):
    def fwd(self, x):
        self.in_hw = tuple(x.shape[1:3])
        y, self.arg = T.maxpool3x3s2_fwd_train(x)
        return y

    def bwd(self, dy):
        return T.maxpool3x3s2_bwd(self.arg, dy.contiguous(), self.in_hw)


class SepBN(object):
    def __init__(self, v, name, dil=1, relu_in=True, relu_out=False):
        self.name = name
        self.dw = Depthwise(name, v(name + "/depthwise_kernel"), dil, relu_in)
        self.pw = Conv(name, v(name + "/pointwise_kernel"))
        self.bn = BatchNorm(name + "_bn", v(name + "_bn/gamma"), v(name + "_bn/beta"), relu_out)

    def fwd(self, x):
        return self.bn.fwd(self.pw.fwd(self.dw.fwd(x)))

    def bwd(self, dy, grads):
        return self.dw.bwd(self.pw.bwd(self.bn.bwd(dy, grads), grads, leaf="pointwise_kernel"), grads)


class ConvBN(object):
    def __init__(self, v, conv_name, bn_name, stride, padding, relu, need_dgrad=True):
        self.conv = Conv(conv_name, v(conv_name + "/kernel"), stride, padding, need_dgrad)
        self.bn = BatchNorm(bn_name, v(bn_name + "/gamma"), v(bn_name + "/beta"), relu)

    def fwd(self, x):
        return self.bn.fwd(self.conv.fwd(x))

    def bwd(self, dy, grads):
        return self.conv.bwd(self.bn.bwd(dy, grads), grads)


class XceptionBodyTraining(object):
    """fwd(images fp32 NCHW) -> (mid [N,h,w,728], out [N,h,w,2048]) NHWC bf16; bwd(d_mid, d_out) -> {name: grad}.
    ``variables``: {TF variable name without the model scope: fp32 CUDA tensor in TF layout}."""

    def __init__(self, variables):
        v = variables.__getitem__
        self.b1c1 = ConvBN(v, "block1_conv1", "block1_conv1_bn", 2, "VALID", True, need_dgrad=False)
        self.b1c2 = ConvBN(v, "block1_conv2", "block1_conv2_bn", 1, "VALID", True)
        self.entry = []
        for blk, idx, first_relu in ((2, 1, False), (3, 2, True), (4, 3, True)):
            self.entry.append((ConvBN(v, "conv2d_%d" % idx, "batch_normalization_%d" % idx, 2, "SAME", False),
                               SepBN(v, "block%d_sepconv1" % blk, relu_in=first_relu),
                               SepBN(v, "block%d_sepconv2" % blk), MaxPool()))
        self.middle = [[SepBN(v, "block%d_sepconv%d" % (i + 5, j)) for j in (1, 2, 3)] for i in range(8)]
        self.exit_res = ConvBN(v, "conv2d_4", "batch_normalization_4", 1, "SAME", False)
        self.b13 = [SepBN(v, "block13_sepconv1"), SepBN(v, "block13_sepconv2")]
        self.b14 = [SepBN(v, "block14_sepconv1", dil=2, relu_in=False, relu_out=True),
                    SepBN(v, "block14_sepconv2", dil=2, relu_in=False, relu_out=True)]

    def fwd(self, images):
        x = _image_nhwc(images)   # 3 channels in rows of 8 (generic strided conv)
        x = self.b1c2.fwd(self.b1c1.fwd(x))
        for res, s1, s2, pool in self.entry:
            x = pool.fwd(s2.fwd(s1.fwd(x))) + res.fwd(x)
        for blk in self.middle:
            y = x
            for s in blk:
                y = s.fwd(y)
            x = x + y
        self.pre_mid = x
        mid = torch.relu(x)
        y = self.b13[1].fwd(self.b13[0].fwd(x)) + self.exit_res.fwd(x)
        return mid, self.b14[1].fwd(self.b14[0].fwd(y))

    def bwd(self, d_mid, d_out, stage_done=None):
        """``stage_done(name)`` is called when every gradient of the 'exit' / 'middle' flow has been written (the
        trainer starts that bucket's all-reduce there)."""
        grads = {}
        d = self.b14[0].bwd(self.b14[1].bwd(d_out, grads), grads)
        dx = self.b13[0].bwd(self.b13[1].bwd(d, grads), grads) + self.exit_res.bwd(d, grads)
        if stage_done is not None:
            stage_done("exit")
        dx = dx + T.relu_bwd(d_mid.contiguous(), self.pre_mid)
        for blk in reversed(self.middle):
            d = dx
            for s in reversed(blk):
                d = s.bwd(d, grads)
            dx = dx + d
        if stage_done is not None:
            stage_done("middle")
        for res, s1, s2, pool in reversed(self.entry):
            dx = s1.bwd(s2.bwd(pool.bwd(dx), grads), grads) + res.bwd(dx, grads)
        self.b1c1.bwd(self.b1c2.bwd(dx, grads), grads)
        return grads


# ---------------------------------------------------------------------------------------------------------------
# The same backbone in the form LightHeadTrainer needs (light_head_rfcn_train.py): gradients accumulate into views of
# the ONE flat all-reduce buffer, every variable owns its momentum slot, the forward is split where the training
# step forks its second stream (the RPN feature exists after the middle flow; the exit flow runs beside the RPN
# losses / proposals / RoI targets).
# ---------------------------------------------------------------------------------------------------------------
class TrainableXceptionBody(XceptionBodyTraining):
    """``XceptionBodyTraining`` over the trainer's parameter classes.  ``store_vars``: {name: fp32 master in TF
    layout} (updated in place by ``update``); ``moving``: {name + '/moving_mean' | '/moving_variance': tensor};
    ``reg``: the trainer's _Registry (gradient views are carved from its flat buffer at ``reg.finalize()``)."""

    def __init__(self, store_vars, moving, reg, conv_params_cls, vec_param_cls, key_prefix="", wg_stream=None,
                 arena=None):
        self.convs, self.vecs = [], []
        self._reg, self._conv_cls, self._vec_cls, self._moving = reg, conv_params_cls, vec_param_cls, moving
        self._vars = store_vars
        XceptionBodyTraining.__init__(self, store_vars)
        # positions in the registry where the middle and exit flows start (their gradients complete, in backward
        # order, before the entry flow's: the trainer all-reduces them as separate buckets)
        self.req_marks = {}
        first_middle, first_exit = self.middle[0][0].dw, self.exit_res.conv
        for layer in self._layers():
            if layer is first_middle:
                self.req_marks["middle"] = len(reg.requests)
            if layer is first_exit:
                self.req_marks["exit"] = len(reg.requests)
            if isinstance(layer, Conv):
                leaf = "pointwise_kernel" if (layer.name + "/pointwise_kernel") in store_vars else "kernel"
                key = layer.name + "/" + leaf
                layer.p = conv_params_cls(reg, [(key_prefix + key, store_vars[key], 0, 0)], layer.kh, layer.kw, layer.cin, layer.cout,
                                          need_dgrad=layer.dpack is not None)
                layer.pack, layer.dpack = layer.p.pack, layer.p.dpack     # the packs the optimizer refreshes
                layer.wg_stream = wg_stream
                layer.arena = arena
                self.convs.append(layer.p)
            elif isinstance(layer, Depthwise):
                layer.master = store_vars[layer.name + "/depthwise_kernel"]
                layer.w9 = layer.master.view(9, layer.C)                   # a VIEW: follows the optimizer's updates
                layer.vec = vec_param_cls(reg, [layer.master], decayed=True)
                self.vecs.append(layer.vec)
            elif isinstance(layer, BatchNorm):
                layer.vec = vec_param_cls(reg, [layer.beta, layer.gamma], decayed=False)
                layer.moving = (moving[layer.name + "/moving_mean"], moving[layer.name + "/moving_variance"])
                self.vecs.append(layer.vec)

    def _layers(self):
        blocks = [self.b1c1, self.b1c2] + [b for e in self.entry for b in e[:3]] + [s for m in self.middle for s in m]
        blocks += [self.exit_res] + self.b13 + self.b14
        for b in blocks:
            if isinstance(b, ConvBN):
                yield b.conv
                yield b.bn
            else:
                yield b.dw
                yield b.pw
                yield b.bn

    # ---- forward, split at the RPN feature -------------------------------------------------------------------
    def fwd_mid(self, images):
        x = _image_nhwc(images)
        x = self.b1c2.fwd(self.b1c1.fwd(x))
        for res, s1, s2, pool in self.entry:
            x = pool.fwd(s2.fwd(s1.fwd(x))) + res.fwd(x)
        for blk in self.middle:
            y = x
            for s in blk:
                y = s.fwd(y)
            x = x + y
        self.pre_mid = x
        return torch.relu(x)

    def fwd_exit(self):
        x = self.pre_mid
        y = self.b13[1].fwd(self.b13[0].fwd(x)) + self.exit_res.fwd(x)
        return self.b14[1].fwd(self.b14[0].fwd(y))

    def fwd(self, images):
        mid = self.fwd_mid(images)
        return mid, self.fwd_exit()

    def update(self, lr, momentum, weight_decay, grad_scale):
        """One momentum-SGD step on every variable of the backbone (L2 on everything but the batch-norm pairs, as
        light_head_rfcn_train.py:420 excludes 'batch_normalization' / '_bn')."""
        for c in self.convs:
            c.update(lr, momentum, weight_decay, grad_scale)
        for v in self.vecs:
            v.update(lr, momentum, weight_decay, grad_scale)


def variable_specs(in_channels=3):
    """(kind, name, shape) of every variable of XceptionBody in the reference's creation order
    (net/xception_body.py:243-376): kind = 'kernel' | 'depthwise_kernel' | 'pointwise_kernel' | 'bn' (shape = channels)."""
    out = []

    def conv(name, k, cin, cout, bn_name):
        out.append(("kernel", name, (k, k, cin, cout)))
        out.append(("bn", bn_name, cout))

    def sep(name, cin, cout):
        out.append(("depthwise_kernel", name, (3, 3, cin, 1)))
        out.append(("pointwise_kernel", name, (1, 1, cin, cout)))
        out.append(("bn", name + "_bn", cout))

    conv("block1_conv1", 3, in_channels, 32, "block1_conv1_bn")
    conv("block1_conv2", 3, 32, 64, "block1_conv2_bn")
    cin = 64
    for blk, idx, filters in ((2, 1, 128), (3, 2, 256), (4, 3, 728)):
        conv("conv2d_%d" % idx, 1, cin, filters, "batch_normalization_%d" % idx)
        sep("block%d_sepconv1" % blk, cin, filters)
        sep("block%d_sepconv2" % blk, filters, filters)
        cin = filters
    for i in range(8):
        for j in (1, 2, 3):
            sep("block%d_sepconv%d" % (i + 5, j), 728, 728)
    conv("conv2d_4", 1, 728, 1024, "batch_normalization_4")
    sep("block13_sepconv1", 728, 728)
    sep("block13_sepconv2", 728, 1024)
    sep("block14_sepconv1", 1024, 1536)
    sep("block14_sepconv2", 1536, 2048)
    return out
