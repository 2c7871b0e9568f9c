"""TEST INFRASTRUCTURE ONLY -- the training-mode XceptionBody (net/xception_body.py:220-379, is_training=True) written
as an explicit forward TAPE and a hand-written layer-by-layer BACKWARD, in PyTorch-CPU float64 without autograd.

Why it exists: the product's training step launches hand-written gradient kernels in a hand-written order (no
autograd on the device); for the ResNet-50 composition that order lives in light_head_rfcn_train.py and is checked
on the GPU against oracle/net_train.py.  The Xception backbone's training path (SURVEY 8 row a15 for the reference's
own backbone) is the next thing to build; this file fixes its op sequence first, on the CPU, where it can be checked
now: every ``bwd`` below is the arithmetic ONE kernel will do --

  Conv.bwd          conv2d_wgrad + conv2d_dgrad (strided / 'valid' / 'same' geometries of the entry flow)
  Depthwise.bwd     depthwise weight gradient (csrc/staged/depthwise_wgrad.cu) + input gradient AS THE FORWARD KERNEL
                    WITH FLIPPED TAPS + relu_bwd when the block ReLUs its input
  BatchNorm.bwd     bn_relu_bwd with / without the ReLU mask: dbeta, dgamma, dx from batch statistics
  MaxPool.bwd       scatter through the recorded arg-max (maxpool3x3s2_bwd) with the residual branch's gradient added

-- and tests/test_xception_backward.py holds the whole thing to torch autograd over oracle/net.xception_body (which is
itself pinned by the reference's own XceptionBody run in training mode, tests/test_netgraph_golden.py /
test_trainstep_golden.py)."""
import torch
import torch.nn.functional as F

from . import net as onet

EPS = onet.EPS_XCEPTION
# True: round every activation / activation-gradient that crosses a kernel boundary to bf16 (what the CUDA twin,
# x-detector_b200/net/xception_train.py, stores between launches); weight gradients stay unrounded (fp32 on
# the device).  Used to CALIBRATE the tolerances of the device-vs-blueprint test, not by the autograd check.
EMULATE_BF16 = False


def _r(t):
    return t.to(torch.bfloat16).to(t.dtype) if EMULATE_BF16 else t


class Conv(object):
    """tf.layers.conv2d, no bias.  kernel [kh,kw,cin,cout]."""

    def __init__(self, name, w, stride=1, padding="SAME"):
        self.name, self.w, self.stride, self.padding = name, w, stride, padding

    def fwd(self, x):
        w = self.w.permute(3, 2, 0, 1).contiguous()
        kh, kw = w.shape[2:]
        self.pads = (0, 0, 0, 0)
        if self.padding == "SAME":
            pt, pb = onet.same_pad(x.shape[2], kh, self.stride)
            pl, pr = onet.same_pad(x.shape[3], kw, self.stride)
            self.pads = (pl, pr, pt, pb)
        self.xp = F.pad(x, self.pads)
        self.wt = w
        return _r(F.conv2d(self.xp, w, stride=self.stride))

    def bwd(self, dy, grads):
        dw = torch.nn.grad.conv2d_weight(self.xp, self.wt.shape, dy, stride=self.stride)
        grads[self.name + "/kernel"] = dw.permute(2, 3, 1, 0).contiguous()
        dxp = torch.nn.grad.conv2d_input(self.xp.shape, self.wt, dy, stride=self.stride)
        pl, pr, pt, pb = self.pads
        return _r(dxp[:, :, pt:dxp.shape[2] - pb, pl:dxp.shape[3] - pr])


def depthwise_fwd(a, taps, dil):
    """Depthwise 3x3 'same' stride-1 correlation; taps [3,3,C]."""
    c = a.shape[1]
    ap = F.pad(a, (dil, dil, dil, dil))
    return F.conv2d(ap, taps.permute(2, 0, 1).reshape(c, 1, 3, 3), dilation=dil, groups=c)


class Depthwise(object):
    """The depthwise stage of tf.layers.separable_conv2d (depth multiplier 1); relu_in = the ReLU that
    relu_separable_bn_block applies first (fused into the kernel's load)."""

    def __init__(self, name, w, dil=1, relu_in=True):
        self.name, self.taps, self.dil, self.relu_in = name, w[..., 0], dil, relu_in   # [3,3,C]

    def fwd(self, x):
        self.x = x
        self.a = torch.relu(x) if self.relu_in else x
        return _r(depthwise_fwd(self.a, self.taps, self.dil))

    def bwd(self, dy, grads):
        d, a = self.dil, self.a
        H, W = a.shape[2:]
        ap = F.pad(a, (d, d, d, d))
        dw = torch.zeros_like(self.taps)
        for kh in range(3):
            for kw in range(3):   # dW[kh,kw,c] = sum_{n,y,x} a(n, y+(kh-1)d, x+(kw-1)d, c) * dy(n,y,x,c)
                dw[kh, kw] = (ap[:, :, kh * d:kh * d + H, kw * d:kw * d + W] * dy).sum(dim=(0, 2, 3))
        grads[self.name + "/depthwise_kernel"] = dw.unsqueeze(-1)
        da = _r(depthwise_fwd(dy, self.taps.flip(0, 1), d))   # the forward kernel on dy with the taps flipped
        return da * (self.x > 0).to(da.dtype) if self.relu_in else da


class BatchNorm(object):
    """tf.layers.batch_normalization(training=True): batch mean / biased variance; optional fused ReLU output."""

    def __init__(self, name, gamma, beta, relu=False):
        self.name, self.gamma, self.beta, self.relu = name, gamma, beta, relu

    def fwd(self, x):
        sh = (1, -1, 1, 1)
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        self.invstd = 1.0 / torch.sqrt(var + EPS)
        self.xhat = (x - mean.view(sh)) * self.invstd.view(sh)
        self.y = self.xhat * self.gamma.view(sh) + self.beta.view(sh)
        return _r(torch.relu(self.y) if self.relu else self.y)

    def bwd(self, dy, grads):
        sh = (1, -1, 1, 1)
        if self.relu:
            dy = dy * (self.y > 0).to(dy.dtype)
        m = dy.shape[0] * dy.shape[2] * dy.shape[3]
        dbeta = dy.sum(dim=(0, 2, 3))
        dgamma = (dy * self.xhat).sum(dim=(0, 2, 3))
        grads[self.name + "/beta"], grads[self.name + "/gamma"] = dbeta, dgamma
        return _r((self.gamma * self.invstd).view(sh) / m * (m * dy - dbeta.view(sh) - self.xhat * dgamma.view(sh)))


class MaxPool(object):
    """tf.layers.max_pooling2d(3, 2, 'same'); the arg-max is recorded by the forward (first maximum wins)."""

    def fwd(self, x):
        pt, pb = onet.same_pad(x.shape[2], 3, 2)
        pl, pr = onet.same_pad(x.shape[3], 3, 2)
        self.pads = (pl, pr, pt, pb)
        xp = F.pad(x, self.pads, value=float("-inf"))
        self.padded_shape = xp.shape
        y, self.arg = F.max_pool2d(xp, 3, 2, return_indices=True)
        return y

    def bwd(self, dy):
        n, c, hp, wp = self.padded_shape
        dxp = torch.zeros((n, c, hp * wp), dtype=dy.dtype)
        dxp.scatter_add_(2, self.arg.reshape(n, c, -1), dy.reshape(n, c, -1))
        dxp = dxp.view(n, c, hp, wp)
        pl, pr, pt, pb = self.pads
        return dxp[:, :, pt:hp - pb, pl:wp - pr]


class SepBN(object):
    """[ReLU ->] depthwise -> pointwise -> batch norm [-> ReLU]  (relu_separable_bn_block :220-234 when relu_in)."""

    def __init__(self, v, name, dil=1, relu_in=True, relu_out=False):
        self.dw = Depthwise(name, v(name + "/depthwise_kernel"), dil, relu_in)
        self.pw = Conv(name, v(name + "/pointwise_kernel"))
        self.pw_name = name
        self.bn = BatchNorm(name + "_bn", v(name + "_bn/gamma"), v(name + "_bn/beta"), relu_out)

    def fwd(self, x):
        return self.bn.fwd(self.pw.fwd(self.dw.fwd(x)))

    def bwd(self, dy, grads):
        d = self.bn.bwd(dy, grads)
        g = {}
        d = self.pw.bwd(d, g)
        grads[self.pw_name + "/pointwise_kernel"] = g[self.pw_name + "/kernel"]
        return self.dw.bwd(d, grads)


class ConvBN(object):
    def __init__(self, v, conv_name, bn_name, stride, padding, relu):
        self.conv = Conv(conv_name, v(conv_name + "/kernel"), stride, padding)
        self.bn = BatchNorm(bn_name, v(bn_name + "/gamma"), v(bn_name + "/beta"), relu)

    def fwd(self, x):
        return self.bn.fwd(self.conv.fwd(x))

    def bwd(self, dy, grads):
        return self.conv.bwd(self.bn.bwd(dy, grads), grads)


class XceptionBodyTape(object):
    """fwd(x) -> (mid, out) as oracle/net.xception_body under BN_TRAINING; bwd(d_mid, d_out) -> (dx, {name: grad}).
    ``variables``: {TF variable name (without the model scope): float64 tensor}."""

    def __init__(self, variables):
        v = variables.__getitem__
        self.b1c1 = ConvBN(v, "block1_conv1", "block1_conv1_bn", 2, "VALID", True)
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

    def fwd(self, x):
        x = self.b1c2.fwd(self.b1c1.fwd(_r(x)))
        for res, s1, s2, pool in self.entry:
            x = _r(pool.fwd(s2.fwd(s1.fwd(x))) + res.fwd(x))
        for blk in self.middle:
            y = x
            for s in blk:
                y = s.fwd(y)
            x = _r(x + y)
        self.pre_mid = x
        mid = torch.relu(x)
        y = _r(self.b13[1].fwd(self.b13[0].fwd(x)) + self.exit_res.fwd(x))
        return mid, self.b14[1].fwd(self.b14[0].fwd(y))

    def bwd(self, d_mid, d_out):
        grads = {}
        d = self.b14[0].bwd(self.b14[1].bwd(d_out, grads), grads)            # gradient of (block13 + residual)
        dx = _r(self.b13[0].bwd(self.b13[1].bwd(d, grads), grads) + self.exit_res.bwd(d, grads))
        dx = _r(dx + d_mid * (self.pre_mid > 0).to(d_mid.dtype))              # the RPN feature is relu(x)
        for blk in reversed(self.middle):
            d = dx
            for s in reversed(blk):
                d = s.bwd(d, grads)
            dx = _r(dx + d)                                                   # identity shortcut
        for res, s1, s2, pool in reversed(self.entry):
            dx = _r(s1.bwd(s2.bwd(pool.bwd(dx), grads), grads) + res.bwd(dx, grads))
        return self.b1c1.bwd(self.b1c2.bwd(dx, grads), grads), grads
