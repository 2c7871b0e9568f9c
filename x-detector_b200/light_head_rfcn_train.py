"""Light-Head R-CNN training step -- the surface of the reference's ``light_head_rfcn_train.py``.

* ``FLAGS`` / ``make_params`` carry the reference's flag names and defaults (light_head_rfcn_train.py:38-170).
* ``LightHeadTrainer.step(images, gt_boxes, gt_labels, keys)`` is one ``train_op`` of ``lighr_head_model_fn``
  (:277-451): forward in training mode (batch-norm with batch statistics), RPN losses on sampled anchors
  (:321-378), proposals (K=10000 -> NMS 0.7 -> 1800) + ``ext_encode_rois`` (64 RoIs/image @ 25 % fg), PsRoIAlign +
  fc head with OHEM (top-32 of the per-RoI loss, get_head net/xception_body.py:504-560), L2 on the non-BN
  variables (:420), backward through everything, momentum-SGD with the piecewise learning rate (:426-441).
  Backbones: ``backbone='resnet50'`` = the ResNet-50 v2 light-head composition of BASELINE configs 2/4;
  ``backbone='xception'`` = the reference's own training backbone (XceptionBody, :289; net/xception_train.py).
* Data parallel: images shard across ranks; the only exchange is the all-reduce (sum) of the flat fp32 gradient
  buffer (``torch.distributed``: NCCL on GPUs), divided by the world size inside the optimizer kernel.  The buffer is
  laid out in forward order, so its tail is complete first in the backward: it is cut into BUCKETS by network stage
  and each bucket's all-reduce is launched on a communication stream as soon as the last gradient of the stage has
  been written, overlapping the rest of the backward (SURVEY 8 e).  Batch-norm statistics stay local to each rank,
  exactly like TF tower replication.

TensorFlow's autodiff is replaced by an explicit backward: every forward layer object below saves what its
gradient needs and ``bwd`` launches the gradient kernels (``ops.conv2d_wgrad`` / ``ops.conv2d_dgrad`` on the
tensor cores, ``ops.train.*`` for batch-norm, pooling, losses).  ``tf.random_shuffle`` calls (select_samples,
_upsample_rois, ext_encode_rois) consume injected uniform key arrays so that the CPU oracle can be driven
identically.  One reference quirk is kept on purpose: with OHEM, ``tf.gather(psroipooled_rois, select_indices,
axis=1)`` (net/xception_body.py:533) applies EVERY image's top-k index row to EVERY image, so the head runs on
N*N*k RoI rows and the loss is their mean.
"""
import math
import os
import types

import torch

from . import _native, ops
from .net.variables import VariableStore
from .ops import conv as conv_ops
from .ops import train as T
from .preprocessing import anchor_manipulator

# flag name -> default (light_head_rfcn_train.py:38-170)
_DEFAULTS = dict(
    num_classes=21, batch_size=8, data_format='channels_last', train_image_size=480, resnet_size=50,
    match_threshold=0.53, neg_threshold_high=0.5, neg_threshold_low=0., fg_ratio=0.25, roi_one_image=64,
    rpn_anchors_per_image=256, rpn_pre_nms_top_n=10000, rpn_post_nms_top_n=1800, rpn_min_size=16 * 1. / 480,
    rpn_nms_thres=0.7, rpn_fg_ratio=0.5, rpn_match_threshold=0.7, rpn_neg_threshold=0.3, using_ohem=True,
    ohem_roi_one_image=32, weight_decay=0.0002, momentum=0.9, learning_rate=1e-3, end_learning_rate=1e-4,
    decay_boundaries='60000, 80000', lr_decay_factors='1, 0.8, 0.1', model_scope='xception_lighthead',
    # data / summary / checkpoint flags: kept with the reference's defaults (the synthetic-tensor launcher ignores
    # the data ones; the checkpoint ones feed utility/train_helper.py)
    num_readers=16, num_preprocessing_threads=48, num_cpu_threads=0, gpu_memory_fraction=1.0,
    data_dir='../PASCAL/VOC_TF/VOC0712TF/', dataset_name='pascalvoc_0712', dataset_split_name='train',
    model_dir='./logs_light/', log_every_n_steps=10, save_summary_steps=500, save_checkpoints_secs=7200,
    train_epochs=None, nms_threshold=0.3, decay_steps=1000, learning_rate_decay_factor=0.96,
    checkpoint_path='./model/xception', checkpoint_model_scope='',
    checkpoint_exclude_scopes='xception_lighthead/rpn_head, xception_lighthead/large_sep_feature, xception_lighthead/final_head',
    ignore_missing_vars=True, run_on_cloud=True, cloud_checkpoint_path='xception_model/xception_model.ckpt',
    # not a reference flag: which backbone builder to use
    backbone='resnet50',
    # not a reference flag: blocks per block_layer of the ResNet v2 composition (resnet_size 50 = 3,4,6,3)
    resnet_layers=(3, 4, 6, 3),
    # not a reference flag: 'bf16' = the throughput path (bf16 activations / operands, fp32 accumulate and masters);
    # 'f16x2' = the fp32-ACCURATE training mode (fp32 activations and gradients, split-operand tensor-core
    # convolutions for forward / input gradient / weight gradient, fp32 batch-norm / pooling kernels): the mode in
    # which the explicit backward is held to float64 autograd (tests/test_train_step_gpu.py)
    precision='bf16',
)
# fp32-accurate mode: the backward runs on gradients scaled by 2^12 (exact) so that they sit inside fp16's range when
# they are split into (hi, lo) planes; the flat gradient buffer is scaled back once at the end
XCEPTION_BODY_WGRAD_STREAM = os.environ.get("XDET_XBODY_WG", "1") == "1"   # (measured both ways, see _build)
WGRAD_STREAM = True    # weight gradients on their own stream beside the input-gradient chain (bf16 precision)
FUSE_BN_STATS = True   # batch statistics from the producing convolution's epilogue (bf16 precision)
_LOSS_SCALE = 4096.0
FLAGS = types.SimpleNamespace(**_DEFAULTS)
pool_method = 'max'  # light_head_rfcn_train.py:199
_BN_DECAY, _BN_EPS = 0.997, 1e-5  # net/resnet_v2.py:37-38


def make_params(**overrides):
    p = dict(_DEFAULTS)
    p.update(overrides)
    return p


def learning_rate(params, global_step):
    """tf.train.piecewise_constant + the end_learning_rate floor (light_head_rfcn_train.py:426-432)."""
    bounds = [int(b) for b in str(params['decay_boundaries']).split(',')]
    factors = [float(f) for f in str(params['lr_decay_factors']).split(',')]
    lr = params['learning_rate'] * factors[sum(1 for b in bounds if global_step > b)]
    return max(lr, params['end_learning_rate'])


# ---------------------------------------------------------------------------------------------------------------
# parameters: fp32 masters under the reference's variable names + momentum + bf16 packs + gradient views
# ---------------------------------------------------------------------------------------------------------------
class _Registry(object):
    """Collects gradient-buffer requests, then carves ONE flat fp32 buffer (the all-reduce payload)."""

    def __init__(self, device):
        self.device, self.requests, self.flat = device, [], None

    def request(self, numel, setter):
        self.requests.append((numel, setter))

    def finalize(self):
        total = sum((n + 3) // 4 * 4 for n, _ in self.requests)
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.device)
        off = 0
        self.offsets = []  # flat offset of request i (and the total at the end): bucket boundaries are taken from here
        for n, setter in self.requests:
            self.offsets.append(off)
            setter(self.flat[off:off + n])
            off += (n + 3) // 4 * 4
        self.offsets.append(off)


class ConvParams(object):
    """One (possibly fused) convolution / dense weight: masters ``[(key, tensor, co_off, ci_off)]`` in TF layout,
    bf16 packs for the forward and input-gradient kernels, packed fp32 gradient ``dw``."""

    def __init__(self, reg, masters, kh, kw, cin, cout, fold=False, need_dgrad=True):
        self.masters, self.kh, self.kw, self.cin, self.cout, self.fold = masters, kh, kw, cin, cout, fold
        dev = masters[0][1].device
        self.mom = [torch.zeros_like(m[1]) for m in masters]
        w = torch.zeros((cout, cin, kh, kw), dtype=torch.float32, device=dev)
        for _, t, co, ci in masters:
            t4 = t if t.dim() == 4 else t.reshape(1, 1, *t.shape)
            w[co:co + t4.shape[3], ci:ci + t4.shape[2]] = t4.permute(3, 2, 0, 1)
        if fold:
            self.pack = ops.pack_fold_weight(w)
            self.cin_pad, shape = 64, (cout, kh, 64)
        else:
            self.pack = ops.pack_conv_weight(w)
            self.cin_pad = (cin + 63) // 64 * 64
            shape = (cout, kh * kw, self.cin_pad)
        self.need_dgrad = need_dgrad
        self.dpack = ops.pack_dgrad_weight(w) if need_dgrad else None
        self.dw = None
        self.precision = conv_ops.PRECISION
        reg.request(shape[0] * shape[1] * shape[2], lambda v: setattr(self, 'dw', v.view(shape)))

    def _dense(self):
        w = torch.zeros((self.cout, self.cin, self.kh, self.kw), dtype=torch.float32, device=self.masters[0][1].device)
        for _, t, co, ci in self.masters:
            t4 = t if t.dim() == 4 else t.reshape(1, 1, *t.shape)
            w[co:co + t4.shape[3], ci:ci + t4.shape[2]] = t4.permute(3, 2, 0, 1)
        return w

    def update(self, lr, momentum, wd, gscale):
        if self.precision == "f16x2":
            # the optimizer kernel refreshes bf16 packs: give it scratch ones, then re-split the updated masters
            if not hasattr(self, "_scratch"):
                cpad = self.cin_pad
                self._scratch = torch.empty((self.cout, (self.kh if self.fold else self.kh * self.kw) * cpad),
                                            dtype=torch.bfloat16, device=self.dw.device)
                copad = (self.cout + 63) // 64 * 64
                self._dscratch = (torch.empty((self.cin, self.kh * self.kw * copad), dtype=torch.bfloat16,
                                              device=self.dw.device) if self.need_dgrad else None)
            for (_, t, co, ci), m in zip(self.masters, self.mom):
                T.sgd_momentum_conv(self.dw, t, m, self._scratch, self._dscratch, lr, momentum, wd, gscale, co, ci, self.fold)
            with conv_ops.precision("f16x2"):
                w = self._dense()
                self.pack = ops.pack_fold_weight(w) if self.fold else ops.pack_conv_weight(w)
                self.dpack = ops.pack_dgrad_weight(w) if self.need_dgrad else None
            return
        for (_, t, co, ci), m in zip(self.masters, self.mom):
            T.sgd_momentum_conv(self.dw, t, m, self.pack, self.dpack, lr, momentum, wd, gscale, co, ci, self.fold)


class VecParam(object):
    """A bias (weight-decayed) or a batch-norm (beta, gamma) pair (not decayed).  The gradient view holds one
    segment of ``seg`` floats per tensor (``seg`` = the channel pitch of the tensor the gradient kernel reduces,
    i.e. the length rounded up to 8): the kernels write it directly, the optimizer reads the leading part."""

    def __init__(self, reg, tensors, decayed):
        self.tensors, self.decayed = tensors, decayed
        self.seg = (tensors[0].numel() + 7) // 8 * 8
        self.mom = [torch.zeros_like(t) for t in tensors]
        self.grad = None
        reg.request(self.seg * len(tensors), lambda v: setattr(self, 'grad', v))

    def update(self, lr, momentum, wd, gscale):
        for i, (t, m) in enumerate(zip(self.tensors, self.mom)):
            T.sgd_momentum_vec(self.grad[i * self.seg:i * self.seg + t.numel()], t, m, lr, momentum,
                               wd if self.decayed else 0.0, gscale)


# ---------------------------------------------------------------------------------------------------------------
# layers with explicit backward
# ---------------------------------------------------------------------------------------------------------------
class _StatsArena(object):
    """Per-step fp32 scratch for the batch statistics the convolutions' epilogues accumulate (conv2d_nhwc(stats=...)): one
    buffer cleared once when the step begins, slots handed out in call order (the first step sizes it)."""

    def __init__(self, device):
        self.device, self.buf, self.pos, self.need = device, None, 0, 0

    def begin(self):
        if self.need and (self.buf is None or self.buf.numel() < self.need):
            self.buf = torch.zeros(self.need, dtype=torch.float32, device=self.device)
        elif self.buf is not None:
            self.buf.zero_()
        self.pos = 0

    def take(self, n):
        if self.buf is not None and self.pos + n <= self.buf.numel():
            v = self.buf[self.pos:self.pos + n]
        else:
            v = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.pos += n
        self.need = max(self.need, self.pos)
        return v


class Conv(object):
    arena = None   # the trainer's _StatsArena (set per layer by LightHeadTrainer._conv)
    wg_stream = None   # set per layer by the trainer: the stream the weight gradients run on

    def __init__(self, params, stride=1, dilation=1, padding="SAME", bias=None):
        self.p, self.stride, self.dil, self.padding, self.bias = params, stride, dilation, padding, bias

    def _geom(self, H, W):
        p, s = self.p, self.stride
        if self.padding == "SAME":
            return "SAME"
        if self.padding == "FIXED":  # conv2d_fixed_padding with strides > 1 (net/resnet_v2.py:89-100)
            pad = (p.kh - 1) // 2
            return (pad, pad, (H + 2 * pad - p.kh) // s + 1, (W + 2 * pad - p.kw) // s + 1)
        return self.padding

    def fwd(self, x, bias_tensor=None, bn_stats=False, **epilogue):
        """``bn_stats``: the output feeds a training-mode batch-norm -- let the kernel's epilogue accumulate its batch
        statistics (bf16 precision; the sums ride on the output tensor as ``_bn_sums`` for BNRelu.fwd)."""
        p = self.p
        self.x = x
        self.in_hw = x.shape[1:3]
        self.geom = self._geom(*self.in_hw)
        b = bias_tensor if bias_tensor is not None else (self.bias.tensors[0] if self.bias is not None else None)
        sums = None
        if bn_stats and FUSE_BN_STATS and self.arena is not None and x.dtype == torch.bfloat16 and p.cout % 8 == 0:
            sums = self.arena.take(2 * p.cout)
            epilogue["stats"] = sums
        self.y = ops.conv2d_nhwc(x, p.pack, p.cout, p.kh, p.kw, dilation=(self.dil, self.dil), padding=self.geom,
                                 strides=(self.stride, self.stride), cin=p.cin, bias=b, **epilogue)
        if sums is not None:
            self.y._bn_sums = sums
        return self.y

    def bwd(self, dy, need_dx=True, dx_residual=None, dx_layout="nhwc_bf16"):
        """dy: NHWC bf16 (channel pitch a multiple of 8, >= cout).  Accumulates dW (and dbias); returns dX."""
        p = self.p

        def weight_side():
            ops.conv2d_wgrad(self.x, dy, p.kh, p.kw, dilation=(self.dil, self.dil), padding=self.geom,
                             strides=(self.stride, self.stride), cin=p.cin, cout=p.cout, dw=p.dw)
            if self.bias is not None:  # dbias = column sums of dy, accumulated straight into the flat gradient buffer
                dy2 = dy.reshape(-1, dy.shape[-1])
                assert dy2.shape[1] == self.bias.seg, (dy2.shape, self.bias.seg)
                T.col_sums_into(dy2, self.bias.grad)

        wg = self.wg_stream
        if wg is not None and need_dx and dy.dtype == torch.bfloat16:
            # nothing downstream of this layer's backward reads dW: the weight gradient runs on its own stream beside
            # the input-gradient chain (joined before the stage's bucket is sent / the optimizer runs)
            wg.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(wg):
                weight_side()
            dy.record_stream(wg)
            wg.xdet_forked = True   # (the trainer joins the stream only if something ran on it this step)
        else:
            weight_side()
        if not need_dx:
            return None
        return ops.conv2d_dgrad(dy, p.dpack, p.cin, p.kh, p.kw, self.in_hw, dilation=(self.dil, self.dil),
                                padding=self.geom, strides=(self.stride, self.stride), cout=p.cout,
                                residual=dx_residual, out_layout=dx_layout)


class BNRelu(object):
    """tf.layers.batch_normalization(training=True) + ReLU (net/resnet_v2.py:41-50) on NHWC bf16.  When the tensor's
    channel pitch exceeds the variable length (490 channels in rows of 496) gamma/beta are zero-extended."""

    def __init__(self, vec, moving, eps=_BN_EPS, decay=_BN_DECAY):
        self.vec, self.moving, self.eps, self.decay = vec, moving, eps, decay  # vec.tensors = (beta, gamma)

    def stats(self, x):
        beta, gamma = self.vec.tensors
        C, cs = beta.numel(), x.shape[-1]
        self.x = x
        if cs == C:
            self.st = T.bn_train(x, gamma, beta, self.eps, self.decay, self.moving[0], self.moving[1])
        else:  # zero tail: statistics of the padding channels are (0, 0) and never reach the variables
            pad = (0, cs - C)
            mm, mv = (torch.nn.functional.pad(t, pad) for t in self.moving)
            self.st = T.bn_train(x, torch.nn.functional.pad(gamma, pad), torch.nn.functional.pad(beta, pad), self.eps,
                                 self.decay, mm, mv)
            self.moving[0].copy_(mm[:C])
            self.moving[1].copy_(mv[:C])
        return self.st

    def fwd(self, x):
        sums = getattr(x, "_bn_sums", None)
        beta, gamma = self.vec.tensors
        if sums is not None and x.dtype == torch.bfloat16 and x.shape[-1] == beta.numel():
            # the producing convolution already summed the batch: finalize + normalise + ReLU in one launch
            self.x = x
            y, self.st = T.bn_train_apply(x, sums, gamma, beta, self.eps, self.decay, self.moving[0], self.moving[1])
            return y
        st = self.stats(x)
        return ops.affine_relu(x, st.scale, st.shift, relu=True)

    def bwd(self, dy, add_in=None):
        cs = self.x.shape[-1]
        assert cs == self.vec.seg and dy.shape == self.x.shape and dy.is_contiguous()
        # the kernel's sums ARE the gradient: [0,cs) = dbeta, [cs,2cs) = dgamma, written into the flat buffer
        # (begin_step cleared the buffer and this is the view's only writer: no second memset)
        return T.bn_relu_bwd_into(dy, self.x, self.st, True, self.vec.grad, add_in, sums_zeroed=True)


class Bottleneck(object):
    """Pre-activation bottleneck (net/resnet_v2.py:142-184; dilated 3x3 as net/xdet_body.py:39-81)."""

    def __init__(self, bn1, proj, c1, bn2, c2, bn3, c3):
        self.bn1, self.proj, self.c1, self.bn2, self.c2, self.bn3, self.c3 = bn1, proj, c1, bn2, c2, bn3, c3

    def fwd(self, x):
        a = self.bn1.fwd(x)
        sc = self.proj.fwd(a) if self.proj is not None else x
        b = self.bn2.fwd(self.c1.fwd(a, bn_stats=True))
        c = self.bn3.fwd(self.c2.fwd(b, bn_stats=True))
        return self.c3.fwd(c, residual=sc, bn_stats=True)   # (the next block's / stage's batch-norm reads the sum)

    def bwd(self, dy, extra_dx=None):
        """dy: gradient of the block output.  Returns the gradient of the block input (+ ``extra_dx``)."""
        dc = self.c3.bwd(dy)
        db = self.c2.bwd(self.bn3.bwd(dc))
        dc1 = self.bn2.bwd(db)
        if self.proj is not None:
            da_p = self.proj.bwd(dy)
            da = self.c1.bwd(dc1, dx_residual=da_p)
            return self.bn1.bwd(da, add_in=extra_dx)
        da = self.c1.bwd(dc1)
        assert extra_dx is None
        return self.bn1.bwd(da, add_in=dy)  # identity shortcut


def allreduce_gradients(flat, group=None):
    """The exchange of a data-parallel step: sum (a bucket of) the flat fp32 gradient buffer over the ranks (NCCL on
    GPUs, gloo in the CPU tests).  Returns the world size; the optimizer kernels divide by it (grad_scale = 1/world)."""
    if not (torch.distributed.is_available() and torch.distributed.is_initialized()):
        return 1
    world = torch.distributed.get_world_size(group)
    if world > 1:
        torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM, group=group)
    return world


def shard_batch(global_batch, world, rank):
    """Images partition across ranks: rank r owns [r*B/world, (r+1)*B/world)."""
    if global_batch % world:
        raise ValueError("global batch %d does not divide over %d ranks" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


# ---------------------------------------------------------------------------------------------------------------
class LightHeadTrainer(object):
    def __init__(self, params=None, seed=0, device="cuda", state_dict=None, process_group=None):
        self.params = p = params or make_params()
        if p['backbone'] not in ('resnet50', 'xception'):
            raise ValueError("backbone must be 'resnet50' or 'xception'")
        self.xception = p['backbone'] == 'xception'
        self.precision = p.get('precision', 'bf16')
        if self.precision not in ('bf16', 'f16x2'):
            raise ValueError("precision must be 'bf16' or 'f16x2'")
        self.f32 = self.precision == 'f16x2'
        self.device = torch.device(device)
        self.store = store = VariableStore(device=device, seed=seed, state_dict=state_dict)
        self.pg = process_group
        self.global_step = 0
        self.reg = reg = _Registry(self.device)
        self.convs, self.vecs = [], []
        self._sgd_plan = None
        self.arena = _StatsArena(self.device)
        self.wg_stream = torch.cuda.Stream(device=self.device) if (WGRAD_STREAM and not self.f32) else None
        size = p['train_image_size']
        self.fmap = size // 16
        creator = anchor_manipulator.AnchorCreator([size] * 2, layers_shapes=[(self.fmap, self.fmap)],
                                                   anchor_scales=[[0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]],
                                                   extra_anchor_scales=[[0.1]], anchor_ratios=[[1., 2., .5]],
                                                   layer_steps=[16])
        all_anchors, num_anchors_list = creator.get_all_anchors()
        self.A = num_anchors_list[0]
        self.enc = anchor_manipulator.AnchorEncoder(all_anchors, num_classes=p['num_classes'], allowed_borders=[0.],
                                                    positive_threshold=p['rpn_match_threshold'],
                                                    ignore_threshold=p['rpn_neg_threshold'],
                                                    prior_scaling=[1., 1., 1., 1.], device=device)
        # anchors in centre form [A_tot,4] and point form (center2point, preprocessing/anchor_manipulator.py:111-112)
        yref, xref, href, wref = self.enc.device_anchors(0)
        fm, A = self.fmap, self.A
        cy = yref.reshape(fm * fm, 1).expand(fm * fm, A).reshape(-1)
        cx = xref.reshape(fm * fm, 1).expand(fm * fm, A).reshape(-1)
        hh = href.reshape(1, A).expand(fm * fm, A).reshape(-1)
        ww = wref.reshape(1, A).expand(fm * fm, A).reshape(-1)
        self.anchors_yxhw = torch.stack([cy, cx, hh, ww], -1).contiguous()
        self.anchors_pt = torch.stack([cy - hh / 2., cx - ww / 2., cy + hh / 2., cx + ww / 2.], -1).contiguous()
        self.side = torch.cuda.Stream(device=self.device)
        self.side2 = torch.cuda.Stream(device=self.device)
        self.comm = torch.cuda.Stream(device=self.device)
        self.marks = []  # (stage name, index of its first gradient request): where the all-reduce buckets start
        with conv_ops.precision(self.precision):
            self._build()
        reg.finalize()
        self.grads = reg.flat
        # buckets in flat (= forward) order: [(name, start, end)]; a stage's bucket is complete when the backward has
        # passed the stage's first layer
        starts = [("first", 0)] + [(n, reg.offsets[i]) for n, i in self.marks]
        self.buckets = [(n, a, (starts[j + 1][1] if j + 1 < len(starts) else reg.offsets[-1]))
                        for j, (n, a) in enumerate(starts)]
        self.overlap_allreduce = True
        self.local_only = False  # tests: keep this rank's own gradient (no exchange)
        self._pending = None

    # ---- bucketed gradient all-reduce ---------------------------------------------------------------------------
    def _world(self):
        if self.local_only or not (torch.distributed.is_available() and torch.distributed.is_initialized()):
            return 1
        return torch.distributed.get_world_size(self.pg)

    def _stage_done(self, name):
        """Every gradient of bucket ``name`` (and of all later buckets) has been written on the current stream: start
        its all-reduce on the communication stream, beside the rest of the backward."""
        if self._pending is None or not self.overlap_allreduce:
            return
        for bname, a, b in self.buckets:
            if bname == name and bname in self._pending and b > a:
                self._pending.remove(bname)
                self.comm.wait_stream(torch.cuda.current_stream())
                if self.wg_stream is not None and getattr(self.wg_stream, "xdet_forked", False):
                    self.comm.wait_stream(self.wg_stream)
                with torch.cuda.stream(self.comm):
                    torch.distributed.all_reduce(self.grads[a:b], op=torch.distributed.ReduceOp.SUM, group=self.pg)

    def _allreduce_finish(self):
        """The buckets not sent yet (all of them without overlap), then JOIN the communication stream."""
        world = self._world()
        if world > 1:
            if self.overlap_allreduce and self._pending is not None:
                for bname, a, b in self.buckets:
                    if bname in self._pending:
                        self._stage_done(bname)
                torch.cuda.current_stream().wait_stream(self.comm)
            else:
                allreduce_gradients(self.grads, self.pg)
        self._pending = None
        return world

    # ---- checkpoints (what tf.estimator / tf.train.Saver do around the reference's train_op, :465-505) -------------
    def trainable_variable_names(self):
        """TF's TRAINABLE_VARIABLES collection of the graph: every variable but the batch-norm moving statistics."""
        return [k for k in self.store.vars if not k.rsplit("/", 1)[-1].startswith("moving_")]

    def _momentum_slots(self):
        """{variable name: its Momentum slot} (``<name>/Momentum`` in a checkpoint, tf.train.MomentumOptimizer)."""
        slots = {}
        for cp in self.convs:
            for (key, _, _, _), m in zip(cp.masters, cp.mom):
                slots[key] = m
        owners = []
        for v in self.vecs:
            owners += list(zip(v.tensors, v.mom))
        owners.append((self.sep_b_biases[1], self.sep_b_bias_mom2))
        for key, var in self.store.vars.items():
            if key in slots:
                continue
            for t, m in owners:  # a fused bias owns one tensor; the named variables are views into it
                off = (var.data_ptr() - t.data_ptr()) // 4
                if var.data_ptr() >= t.data_ptr() and off + var.numel() <= t.numel():
                    slots[key] = m.view(-1)[off:off + var.numel()].view(var.shape)
                    break
        return slots

    def save_checkpoint(self, model_dir, keep=5):
        """Write ``model_dir/model.ckpt-<global_step>`` as a TensorFlow V2 checkpoint: every variable under its TF
        name, the Momentum slots and ``global_step``; update the ``checkpoint`` state file.  Returns the prefix."""
        import os

        from .utility import tensor_bundle as tb
        torch.cuda.synchronize(self.device)
        prefix = os.path.join(model_dir, "model.ckpt-%d" % self.global_step)
        w = tb.TensorBundleWriter(prefix)
        for k, v in self.store.vars.items():
            w.add(k, v.detach().cpu().numpy())
        for k, m in self._momentum_slots().items():
            w.add(k + "/Momentum", m.detach().cpu().numpy())
        import numpy as np
        w.add("global_step", np.array(self.global_step, dtype=np.int64))
        w.finish()
        tb.update_checkpoint_state(model_dir, prefix, keep=keep)
        return prefix

    def restore_checkpoint(self, prefix_or_dir):
        """Resume: variables, Momentum slots and global_step from a checkpoint written by ``save_checkpoint`` (or by the
        reference's Estimator for the same graph).  The bf16 packs are rebuilt from the restored masters."""
        from .utility import tensor_bundle as tb
        from .utility import train_helper as th
        path = th.resolve_checkpoint_path(prefix_or_dir)
        if path is None:
            raise FileNotFoundError("no checkpoint under %s" % prefix_or_dir)
        r = tb.TensorBundleReader(path)
        missing = [k for k in self.store.vars if not r.has_tensor(k)]
        if missing:
            raise KeyError("checkpoint %s lacks %d variable(s) of the model, e.g. %s" % (path, len(missing), missing[:3]))
        for k, v in self.store.vars.items():
            v.copy_(torch.from_numpy(r.get_tensor(k)).to(v.dtype).reshape(v.shape))
        for k, m in self._momentum_slots().items():
            if r.has_tensor(k + "/Momentum"):
                m.copy_(torch.from_numpy(r.get_tensor(k + "/Momentum")).to(m.dtype).reshape(m.shape))
        if r.has_tensor("global_step"):
            self.global_step = int(r.get_tensor("global_step"))
        self.refresh_packs()
        return path

    def refresh_packs(self):
        """Re-derive the bf16 forward / input-gradient packs from the fp32 masters (after loading weights): one
        optimizer update with zero learning rate, zero gradient and zero decay rewrites them and changes nothing else."""
        self.grads.zero_()
        saved = [[m.clone() for m in c.mom] for c in self.convs]
        for c in self.convs:
            c.update(0.0, 1.0, 0.0, 1.0)
        for c, ms in zip(self.convs, saved):
            for m, m0 in zip(c.mom, ms):
                m.copy_(m0)

    def comm_info(self):
        return {"collective": "ncclAllReduce(sum, fp32) per bucket on a communication stream, launched when the "
                              "stage's last gradient is written (overlaps the rest of the backward)",
                "overlap": bool(self.overlap_allreduce),
                "buckets": [{"stage": n, "bytes": int((b - a) * 4)} for n, a, b in reversed(self.buckets)]}

    # ---- variable creation in the reference's naming order --------------------------------------------------
    def _conv(self, cin, cout, k, stride=1, dil=1, fold=False, need_dgrad=True, init=None):
        s = self.store
        name = s.auto_name("conv2d")
        with s.scope(name):
            key, t = s.get("kernel", (k, k, cin, cout), init or s.variance_scaling)
        cp = ConvParams(self.reg, [(key, t, 0, 0)], k, k, cin, cout, fold=fold, need_dgrad=need_dgrad)
        self.convs.append(cp)
        padding = "SAME" if stride == 1 else "FIXED"
        c = Conv(cp, stride=stride, dilation=dil, padding=padding)
        c.arena = self.arena
        c.wg_stream = self.wg_stream
        return c

    def _bn(self, channels, name=None):
        s = self.store
        bn = s.batch_norm(name or s.auto_name("batch_normalization"), channels)
        vec = VecParam(self.reg, [bn["beta"][1], bn["gamma"][1]], decayed=False)
        self.vecs.append(vec)
        return BNRelu(vec, (bn["mean"][1], bn["var"][1]))

    def _bias(self, tensors):
        vec = VecParam(self.reg, tensors, decayed=True)
        self.vecs.append(vec)
        return vec

    def _block(self, cin, filters, project, stride, dil):
        bn1 = self._bn(cin)
        proj = self._conv(cin, 4 * filters, 1, stride=stride if dil == 1 else 1) if project else None
        c1 = self._conv(cin, filters, 1)
        bn2 = self._bn(filters)
        c2 = self._conv(filters, filters, 3, stride=stride if dil == 1 else 1, dil=dil)
        bn3 = self._bn(filters)
        c3 = self._conv(filters, 4 * filters, 1)
        return Bottleneck(bn1, proj, c1, bn2, c2, bn3, c3)

    def _build(self):
        s, p = self.store, self.params
        nc, A = p['num_classes'], self.A
        with s.scope(p['model_scope']):
            if self.xception:
                # the reference's own training backbone (light_head_rfcn_train.py:289): variables under the reference's
                # names, gradients / momentum / moving statistics through this trainer's parameter classes
                from .net import xception_train as xt
                vars_, moving = {}, {}
                for kind, name, shape in xt.variable_specs(3):
                    if kind == "bn":
                        bn = s.batch_norm(name, shape)
                        vars_[name + "/gamma"], vars_[name + "/beta"] = bn["gamma"][1], bn["beta"][1]
                        moving[name + "/moving_mean"], moving[name + "/moving_variance"] = bn["mean"][1], bn["var"][1]
                    else:
                        with s.scope(name):
                            vars_[name + "/" + kind] = s.get(kind, shape, s.glorot_normal)[1]
                self.body = xt.TrainableXceptionBody(vars_, moving, self.reg, ConvParams, VecParam,
                                                     key_prefix=p['model_scope'] + "/",
                                                     wg_stream=self.wg_stream if XCEPTION_BODY_WGRAD_STREAM else None,
                                                     arena=self.arena if FUSE_BN_STATS else None)
                # (measured: with the first depthwise weight-gradient kernel -- 163 us per layer -- the pointwise weight
                # gradients beside the chain LOST 7 %; with the row-sliding one they gain 3 %: 740 -> 760 img/s)
                self.convs += self.body.convs
                self.vecs += self.body.vecs
                self.marks += [("middle", self.body.req_marks["middle"]), ("exit", self.body.req_marks["exit"])]
                rpn_cin = 728
            else:
                self.stem = self._conv(3, 64, 7, stride=2, fold=True, need_dgrad=False)
                self.layers = []
                cin = 64
                nb = p['resnet_layers']
                for filters, blocks, stride, dil in ((64, nb[0], 1, 1), (128, nb[1], 2, 1), (256, nb[2], 2, 1),
                                                     (512, nb[3], 2, 2)):
                    if len(self.layers) >= 2:
                        self.marks.append(("layer%d" % (len(self.layers) + 1), len(self.reg.requests)))
                    layer = []
                    for i in range(blocks):
                        layer.append(self._block(cin, filters, i == 0, stride if i == 0 else 1, dil))
                        cin = 4 * filters
                    self.layers.append(layer)
                    if len(self.layers) == 3:
                        self.bn_rpn = self._bn(cin)  # batch_norm_relu -> RPN feature (created before block_layer4)
                        # (it lands in the layer-3 bucket: its gradient is written right before layer 3's)
                self.marks.append(("heads", len(self.reg.requests)))
                self.bn_final = self._bn(cin)
                rpn_cin = 1024
            if self.xception:
                self.marks.append(("heads", len(self.reg.requests)))

            def named_conv(kh, kw, cin_, cout_, init=None):
                name = s.auto_name("conv2d")
                with s.scope(name):
                    k = s.get("kernel", (kh, kw, cin_, cout_), init or s.glorot_normal)
                    b = s.get("bias", (cout_,), s.zeros)
                return k, b

            def fuse_bias(pairs):
                """Re-home separate bias variables as views of ONE tensor (the fused convolution's bias)."""
                fused = torch.cat([t for _, t in pairs]).contiguous()
                off, views = 0, []
                for key, t in pairs:
                    v = fused[off:off + t.numel()]
                    s.vars[key] = v
                    views.append(v)
                    off += t.numel()
                return fused, views

            with s.scope('rpn_head'):
                k0, b0 = named_conv(3, 3, rpn_cin, 512)
                k1, b1 = named_conv(1, 1, 512, 2 * A)
                k2, b2 = named_conv(1, 1, 512, 4 * A)
            cp0 = ConvParams(self.reg, [(k0[0], k0[1], 0, 0)], 3, 3, rpn_cin, 512)
            cp12 = ConvParams(self.reg, [(k1[0], k1[1], 0, 0), (k2[0], k2[1], 2 * A, 0)], 1, 1, 512, 6 * A)
            self.convs += [cp0, cp12]
            self.rpn_conv = Conv(cp0, bias=self._bias([b0[1]]))
            fused12, _ = fuse_bias([b1, b2])
            self.rpn_out = Conv(cp12, bias=self._bias([fused12]))

            with s.scope('large_sep_feature'):
                with s.scope("Branch_0"):
                    a0k, a0b = named_conv(15, 1, 2048, 256)
                    b0k, b0b = named_conv(1, 15, 256, 490)
                with s.scope("Branch_1"):
                    a1k, a1b = named_conv(15, 1, 2048, 256)
                    b1k, b1b = named_conv(1, 15, 256, 490)
                self.bn_sep = self._bn(490, name=s.auto_name("batch_normalization"))
            cpa = ConvParams(self.reg, [(a0k[0], a0k[1], 0, 0), (a1k[0], a1k[1], 256, 0)], 15, 1, 2048, 512)
            cpa.kh, cpa.kw = 15, 1
            cpb = ConvParams(self.reg, [(b0k[0], b0k[1], 0, 0), (b1k[0], b1k[1], 0, 256)], 1, 15, 512, 490)
            self.convs += [cpa, cpb]
            fused_a, _ = fuse_bias([a0b, a1b])
            self.sep_a = Conv(cpa, bias=self._bias([fused_a]))
            # the 1x15 biases of the two branches both receive the summed-output gradient; forward adds their sum
            self.sep_b_biases = (b0b[1], b1b[1])
            self.sep_b_bias_vec = self._bias([b0b[1]])       # gradient view shared by both (identical gradients)
            self.sep_b_bias_mom2 = torch.zeros_like(b1b[1])
            self.sep_b = Conv(cpb, bias=None)

            with s.scope('final_head'):
                with s.scope("subnet_fc"):
                    f1k = s.get("kernel", (490, 2048), s.glorot_normal)
                    f1b = s.get("bias", (2048,), s.zeros)
                with s.scope("fc_cls"):
                    fck = s.get("kernel", (2048, nc), s.glorot_normal)
                    fcb = s.get("bias", (nc,), s.zeros)
                with s.scope("fc_loc"):
                    flk = s.get("kernel", (2048, 4), s.glorot_normal)
                    flb = s.get("bias", (4,), s.zeros)
            cpf1 = ConvParams(self.reg, [(f1k[0], f1k[1], 0, 0)], 1, 1, 490, 2048)
            cpf2 = ConvParams(self.reg, [(fck[0], fck[1], 0, 0), (flk[0], flk[1], nc, 0)], 1, 1, 2048, nc + 4)
            self.convs += [cpf1, cpf2]
            self.fc1 = Conv(cpf1, bias=self._bias([f1b[1]]))
            fused_f2, _ = fuse_bias([fcb, flb])
            self.fc2 = Conv(cpf2, bias=self._bias([fused_f2]))
            for c in (self.rpn_conv, self.rpn_out, self.sep_a, self.sep_b, self.fc1, self.fc2):
                c.wg_stream = self.wg_stream

    # ---- forward pieces ----------------------------------------------------------------------------------------
    def _rows(self, feat, pitch):
        """PsRoIAlign rows for the dense layers: bf16 rows of ``pitch`` channels (fp32 as they are in the f16x2 mode)."""
        return feat.contiguous() if self.f32 else ops.f32_to_bf16_rows(feat.contiguous(), pitch)

    def _head_fwd(self, feat_bf16):
        """feat [M, 496] bf16 (or [M, 490] fp32) -> (h [1,1,M,2048] ReLU'd, out [M, nc+4] fp32)."""
        M = feat_bf16.shape[0]
        h = self.fc1.fwd(feat_bf16.reshape(1, 1, M, -1), relu=True)
        out = self.fc2.fwd(h, out_layout="nhwc_f32")
        return h, out.reshape(M, -1)

    def _head_loss(self, out, labels_i32, targets, w_rows):
        """Per-row loss of head_loss_func (:381-399): CE + smooth-L1 * [label>0] / fg_ratio, and its gradient
        w.r.t. ``out`` with every row weighted by ``w_rows`` (1/M for the mean)."""
        nc = self.params['num_classes']
        pos = (labels_i32 > 0).float() / self.params['fg_ratio']
        dout = torch.zeros_like(out)
        ce, _ = T.softmax_ce(out, labels_i32, nc, w_all=w_rows, dlogits=dout)
        l1, _ = T.smooth_l1(out[:, nc:], targets, row_w=pos, w_all=w_rows, dpred=dout[:, nc:])
        return ce + l1, dout

    # ---- one training step ------------------------------------------------------------------------------------
    def step(self, images, gt_boxes, gt_labels, keys, apply_update=True, inject=None):
        """images [N,3,H,W] fp32; gt_boxes [N,G,4] fp32 (ymin,xmin,ymax,xmax in [0,1]); gt_labels [N,G] int32
        (0 = padding); keys: dict of uniform [0,1) fp32 arrays standing in for tf.random_shuffle:
        'rpn_fg','rpn_bg' [1, N*A_tot], 'rpn_up' [1, N*256], 'prop' [N, 1800], 'roi_fg','roi_bg' [N, 1800+G],
        'roi_up' [N, 64].  ``inject`` (tests): {'rois_all','roi_idx','rpn_idx','ohem_idx'} override the discrete
        selections.  Returns a dict of scalar losses (device tensors) and the selections."""
        with conv_ops.precision(self.precision):
            return self._step(images, gt_boxes, gt_labels, keys, apply_update, inject)

    def _step(self, images, gt_boxes, gt_labels, keys, apply_update, inject):
        """The step = the pieces below in the order (and on the streams) the trainer wants them.  The reference-named
        builders (net/xception_body.py with is_training=True) call the same pieces one by one on ONE stream."""
        self.begin_step(images, gt_boxes, gt_labels, keys, inject)
        t = self.t
        rpn_feat = self.fwd_backbone_mid()
        self.fwd_rpn(rpn_feat)
        # FORK: proposals + RoI targets on a second stream, anchor targets + RPN losses on a third (the reference pins
        # proposals / ext_encode_rois to /cpu:0; here they run beside block_layer4 / the exit flow + large_sep_kernel,
        # whose convolutions leave a few SMs to the one-CTA-per-image kernels meanwhile).  The RPN losses are ready first:
        # the RPN head's backward fills the time the proposals still need, and only then the head waits for its RoIs.
        main = torch.cuda.current_stream()
        self.fwd_rpn_decode()
        self.side.wait_stream(main)
        self.side2.wait_stream(main)
        with torch.cuda.stream(self.side):
            self.fwd_proposals_and_targets()
            ev_rois = torch.cuda.Event()
            ev_rois.record(self.side)
        with torch.cuda.stream(self.side2):
            self.fwd_rpn_losses()
            t.ev_rpn = torch.cuda.Event()
            t.ev_rpn.record(self.side2)
        conv_ops.MAX_CTAS = 148 - 12
        try:
            backbone = self.fwd_backbone_exit()
            self.fwd_thin(backbone)
        finally:
            conv_ops.MAX_CTAS = 0
        self.backward_rpn_head()   # JOIN (RPN-loss stream)
        main.wait_event(ev_rois)   # JOIN (proposal stream)
        for t_ in (t.rois_all, t.rlab, t.rtgt, t.rsc, t.roi_idx, t.rois, t.roi_tgt, t.roi_lab, t.yxhw):
            t_.record_stream(main)
        self.fwd_head()
        self.backward()
        self.apply_gradients(apply_update)
        return self.outputs()

    # ---- the pieces -------------------------------------------------------------------------------------------
    def begin_step(self, images, gt_boxes, gt_labels, keys, inject=None):
        """Start a step: zero the flat gradient buffer, open the tape ``self.t``."""
        p = self.params
        t = self.t = types.SimpleNamespace()
        t.images, t.gt_boxes, t.gt_labels, t.keys, t.inject = images, gt_boxes, gt_labels, keys, (inject or {})
        t.N = images.shape[0]
        # fp32-accurate mode: the backward carries gradients scaled by S (see _LOSS_SCALE); act = activation dtype
        t.S = _LOSS_SCALE if self.f32 else 1.0
        t.act = torch.float32 if self.f32 else torch.bfloat16
        self.grads.zero_()
        self.arena.begin()
        # (fp32-accurate mode: the buffer is rescaled before it is sent, so its buckets go out together at the end)
        self._pending = set(n for n, _, _ in self.buckets) if (self._world() > 1 and not self.f32) else None
        del p

    def fwd_backbone_mid(self):
        """Backbone up to the RPN feature (XceptionBody's middle flow / block_layer1-3 + batch_norm_relu)."""
        t = self.t
        images = t.images
        if self.xception:
            rpn_feat = self.body.fwd_mid(images)
            fm = rpn_feat.shape[1]
            assert fm == self.fmap, "the anchors are laid out for a %dx%d map, the backbone produced %dx%d" % (
                self.fmap, self.fmap, fm, fm)
        else:
            Wimg = images.shape[3]
            Wo = (Wimg + 6 - 7) // 2 + 1
            wp = (max((Wo - 1) * 2 + 8, Wimg + 3) + 7) // 8 * 8
            if self.f32:  # row-padded f16x2 planes of the image (a fresh buffer: the tape keeps it for the weight gradient)
                x8 = ops.split2(images.permute(0, 2, 3, 1), cin=3, pad_w=(wp, 3)).clone()
            else:
                x8 = ops.image_to_nhwc8(images.contiguous(), 3, wp)
            Ho = (images.shape[2] + 6 - 7) // 2 + 1
            self.stem.x = x8
            y0 = ops.conv2d_nhwc(x8, self.stem.p.pack, 64, 7, 7, padding=(3, 3, Ho, Wo), strides=(2, 2), cin=3,
                                 fold_w=(Wimg, 3))
            x, t.pool_arg = T.maxpool3x3s2_fwd_train(y0)
            for li in range(3):
                for blk in self.layers[li]:
                    x = blk.fwd(x)
            t.x3, t.x8, t.y0_hw, t.stem_geom = x, x8, tuple(y0.shape[1:3]), (Ho, Wo, Wimg)
            rpn_feat = self.bn_rpn.fwd(x)
        t.rpn_feat = rpn_feat
        return rpn_feat

    def fwd_rpn(self, rpn_feat):
        """get_rpn (net/xception_body.py:381-400): -> [N,fm,fm,6A] fp32, logits [0,2A), deltas [2A,6A)."""
        t = self.t
        t.r = self.rpn_conv.fwd(rpn_feat, relu=True)
        t.rpn_out = self.rpn_out.fwd(t.r, out_layout="nhwc_f32")
        return t.rpn_out

    def fwd_rpn_decode(self):
        """Objectness (softmax[:, -1]) and decode_all_anchors of the RPN head's output (train:295-319)."""
        t, A = self.t, self.A
        t.score, t.boxes = ops.rpn_decode(t.rpn_out, 0, 2 * A, self.enc.device_anchors(0), A)
        return t.score, t.boxes

    def fwd_rpn_losses(self):
        """Anchor targets, select_samples (train:321-358), the two RPN losses (:361-378) and their gradient with respect
        to the RPN head's output."""
        t, p, A = self.t, self.params, self.A
        N, fm, S, inject, keys = t.N, self.fmap, t.S, t.inject, t.keys
        rpn_out = t.rpn_out
        if not hasattr(t, "score"):
            self.fwd_rpn_decode()
        t.glabels, t.gtargets, _ = T.match_encode(self.anchors_pt, t.gt_boxes, t.gt_labels, 0.0,
                                                  p['rpn_match_threshold'], p['rpn_neg_threshold'],
                                                  ref_yxhw=self.anchors_yxhw)
        n_rpn = N * p['rpn_anchors_per_image']
        exp_fg = int(round(n_rpn * p['rpn_fg_ratio']))
        if 'rpn_idx' in inject:
            rpn_idx = inject['rpn_idx']
        else:
            rpn_idx, _ = T.sample_fg_bg(t.glabels.reshape(1, -1), None, 0.0, exp_fg, n_rpn, keys['rpn_fg'],
                                        keys['rpn_bg'], keys['rpn_up'])
            rpn_idx = rpn_idx.reshape(-1).long()
        cls_all = rpn_out[..., :2 * A].reshape(-1, 2)   # plumbing: [N*A_tot, 2] copies of the two channel groups
        loc_all = rpn_out[..., 2 * A:].reshape(-1, 4)
        s_cls, s_loc = cls_all.index_select(0, rpn_idx), loc_all.index_select(0, rpn_idx)
        s_lab = (t.glabels.reshape(-1).index_select(0, rpn_idx) > 0).to(torch.int32)
        s_tgt = t.gtargets.reshape(-1, 4).index_select(0, rpn_idx).contiguous()
        # (w_all weights the GRADIENT only: the loss rows are unscaled)
        rpn_ce_rows, d_s_cls = T.softmax_ce(s_cls, s_lab, 2, w_all=S / n_rpn)
        posm = s_lab.float()
        npos = posm.sum().clamp(min=1.0)
        row_w = posm / (npos * p['rpn_fg_ratio'])
        rpn_l1_rows, d_s_loc = T.smooth_l1(s_loc, s_tgt, row_w=row_w, w_all=S)
        t.rpn_ce, t.rpn_loc, t.rpn_idx = rpn_ce_rows.mean(), rpn_l1_rows.sum(), rpn_idx
        # gradient of the RPN losses w.r.t. the head output, scattered back to the dense [N,fm,fm,6A] tensor
        d_cls = torch.zeros_like(cls_all).index_add_(0, rpn_idx, d_s_cls)
        d_loc = torch.zeros_like(loc_all).index_add_(0, rpn_idx, d_s_loc)
        cpitch = (6 * A + 7) // 8 * 8
        d_rpn = torch.zeros((N, fm, fm, cpitch), dtype=t.act, device=self.device)
        d_rpn[..., :2 * A] = d_cls.reshape(N, fm, fm, 2 * A)
        d_rpn[..., 2 * A:6 * A] = d_loc.reshape(N, fm, fm, 4 * A)
        t.d_rpn = d_rpn
        return t.rpn_ce, t.rpn_loc

    def fwd_proposals_and_targets(self):
        """get_proposals in training mode (net/xception_body.py:402-448: clip, top-k, NMS, upsample) + its
        ``encode_fn`` = ext_encode_rois (append the ground truth, match, sample roi_one_image RoIs per image)."""
        t, p = self.t, self.params
        N, inject, keys = t.N, t.inject, t.keys
        gt_boxes, gt_labels = t.gt_boxes, t.gt_labels
        if not hasattr(t, "score"):
            self.fwd_rpn_decode()
        if 'rois_all' in inject:
            rois_all = inject['rois_all']
        else:
            props, _, _ = ops.rpn_select(t.score, t.boxes, p['rpn_pre_nms_top_n'], p['rpn_post_nms_top_n'],
                                         p['rpn_nms_thres'], p['rpn_min_size'], keys['prop'])
            rois_all = torch.cat([props, gt_boxes * (gt_labels > 0).unsqueeze(-1).float()], dim=1).contiguous()
        rlab, rtgt, rsc = T.match_encode(rois_all, gt_boxes, gt_labels, 0.1, p['match_threshold'],
                                         p['neg_threshold_high'])
        # the reference appends only the VALID ground-truth boxes (tf.boolean_mask, anchor_manipulator.py:345-347);
        # here the padded slots ride along as zero boxes: mark them 'ignore' so that no threshold setting can ever
        # sample them (with neg_threshold_low < 0 they would qualify as background)
        G = gt_labels.shape[1]
        rlab[:, rlab.shape[1] - G:].masked_fill_(gt_labels <= 0, -1)
        R = p['roi_one_image']
        if 'roi_idx' in inject:
            roi_idx = inject['roi_idx']
        else:
            roi_idx, _ = T.sample_fg_bg(rlab, rsc, p['neg_threshold_low'], int(round(R * p['fg_ratio'])), R,
                                        keys['roi_fg'], keys['roi_bg'], keys['roi_up'])
            roi_idx = roi_idx.long()
        rois = torch.gather(rois_all, 1, roi_idx.unsqueeze(-1).expand(N, R, 4)).contiguous()
        t.roi_tgt = torch.gather(rtgt, 1, roi_idx.unsqueeze(-1).expand(N, R, 4)).contiguous()
        t.roi_lab = torch.gather(rlab, 1, roi_idx).contiguous()
        t.roi_sc = torch.gather(rsc, 1, roi_idx).contiguous()
        h_, w_ = rois[..., 2] - rois[..., 0], rois[..., 3] - rois[..., 1]
        t.yxhw = torch.stack([rois[..., 0] + h_ / 2., rois[..., 1] + w_ / 2., h_, w_], dim=-1).contiguous()  # _point2center
        t.rois_all, t.rlab, t.rtgt, t.rsc, t.roi_idx, t.rois = rois_all, rlab, rtgt, rsc, roi_idx, rois
        return rois, t.roi_tgt, t.roi_lab, t.roi_sc

    def fwd_backbone_exit(self):
        """The rest of the backbone (XceptionBody's exit flow / block_layer4 + batch_norm_relu) -> [N,fm,fm,2048]."""
        t = self.t
        if self.xception:
            t.backbone = self.body.fwd_exit()
        else:
            x = t.x3
            for blk in self.layers[3]:
                x = blk.fwd(x)
            t.backbone = self.bn_final.fwd(x)
        return t.backbone

    def fwd_thin(self, backbone):
        """large_sep_kernel (net/xception_body.py:450-475) with batch statistics -> thin feature map, fp32 NCHW."""
        t = self.t
        N, fm = t.N, self.fmap
        mid = self.sep_a.fwd(backbone)
        bias_b = self.sep_b_biases[0] + self.sep_b_biases[1]
        # 490 channels live in rows of 496 (16-byte pixel strides for TMA and the vector kernels), zero tail
        o_buf = torch.zeros((N, fm, fm, 496), dtype=t.act, device=self.device)
        self.sep_b.fwd(mid, bias_tensor=bias_b, out=o_buf[..., :490])
        st_sep = self.bn_sep.stats(o_buf)
        if self.f32:  # (layout plumbing of the verification mode: NHWC -> the NCHW PsRoIAlign's contract asks for)
            t.thin = torch.relu(o_buf[..., :490] * st_sep.scale[:490] + st_sep.shift[:490]).permute(0, 3, 1, 2).contiguous()
        else:
            t.thin = T.affine_relu_to_nchw_f32(o_buf, st_sep.scale, st_sep.shift, relu=True, C=490)
        return t.thin

    def fwd_head(self):
        """get_head (net/xception_body.py:477-560): PsRoIAlign, fc head, OHEM (no-grad pass, per-RoI loss, top-k,
        the reference's axis-1 gather), the head loss and its gradient with respect to the head's output."""
        t, p, nc = self.t, self.params, self.params['num_classes']
        N, S, inject = t.N, t.S, t.inject
        R = p['roi_one_image']
        t.pooled, t.pindex = ops.ps_roi_align(t.thin, t.yxhw, 7, 7, pool_method)
        feat = t.pooled.reshape(N * R, -1)
        cin = feat.shape[1]
        pitch = (cin + 7) // 8 * 8
        lab_flat, tgt_flat = t.roi_lab.reshape(-1), t.roi_tgt.reshape(-1, 4)
        if p['using_ohem']:
            k = min(p['ohem_roi_one_image'], R)
            if 'ohem_idx' in inject:
                sel = inject['ohem_idx']
            else:
                _, out1 = self._head_fwd(self._rows(feat, pitch))
                loss1, _ = self._head_loss(out1, lab_flat, tgt_flat, 1.0)
                sel = torch.topk(loss1.reshape(N, R), k, dim=1).indices  # selection only (tf.nn.top_k, :529)
            # tf.gather(x, select_indices, axis=1) with a [N,k] index: every image gets every image's rows
            flat_sel = (torch.arange(N, device=self.device).view(N, 1, 1) * R + sel.view(1, N, k)).reshape(-1)
            feat2 = feat.index_select(0, flat_sel)
            lab2 = t.roi_lab.index_select(1, sel.reshape(-1)).reshape(-1).contiguous()
            tgt2 = t.roi_tgt.index_select(1, sel.reshape(-1)).reshape(-1, 4).contiguous()
        else:
            sel, flat_sel, feat2, lab2, tgt2 = None, None, feat, lab_flat, tgt_flat
        M2 = feat2.shape[0]
        a2 = self._rows(feat2, pitch)
        t.h2, t.out2 = self._head_fwd(a2)
        head_rows, t.dout2 = self._head_loss(t.out2, lab2, tgt2, S / M2)
        t.head_loss = head_rows.mean()
        t.sel, t.flat_sel, t.feat, t.M2, t.cin = sel, flat_sel, feat, M2, cin
        return t.out2[:, :nc], t.out2[:, nc:], t.head_loss

    def backward_rpn_head(self):
        """Backward of the RPN head from the gradient of its two losses: needs only ``fwd_rpn_losses`` (not the proposals
        or the head), so the step runs it early; leaves the gradient with respect to the RPN feature in ``t.d_rpn_feat``."""
        t = self.t
        if getattr(t, "ev_rpn", None) is not None:  # the losses were computed on another stream
            main = torch.cuda.current_stream()
            main.wait_event(t.ev_rpn)
            for t_ in (t.glabels, t.gtargets, t.rpn_idx, t.d_rpn, t.rpn_ce, t.rpn_loc):
                t_.record_stream(main)
        dr = T.relu_bwd(self.rpn_out.bwd(t.d_rpn), t.r)
        t.d_rpn_feat = self.rpn_conv.bwd(dr)
        return t.d_rpn_feat

    def backward(self):
        """The explicit backward of everything above (TensorFlow: optimizer.compute_gradients, train:436-441)."""
        t, nc = self.t, self.params['num_classes']
        N, fm = t.N, self.fmap
        # ---- head ----
        dpitch = (nc + 4 + 7) // 8 * 8
        if self.f32:
            d_out_b = torch.nn.functional.pad(t.dout2, (0, dpitch - t.dout2.shape[1])).reshape(1, 1, t.M2, dpitch)
        else:
            d_out_b = ops.f32_to_bf16_rows(t.dout2, dpitch).reshape(1, 1, t.M2, dpitch)
        dh = self.fc2.bwd(d_out_b)
        dh = T.relu_bwd(dh, t.h2)
        dfeat2 = self.fc1.bwd(dh, dx_layout="nhwc_f32").reshape(t.M2, t.cin)
        if t.flat_sel is not None:
            dfeat = torch.zeros_like(t.feat).index_add_(0, t.flat_sel, dfeat2)
        else:
            dfeat = dfeat2
        d_thin = ops.ps_roi_align_grad(t.thin, t.yxhw, dfeat.reshape(t.pooled.shape).contiguous(), t.pindex, 7, 7,
                                       pool_method)
        # ---- thin feature map ----
        if self.f32:
            d_o = torch.nn.functional.pad(d_thin.permute(0, 2, 3, 1), (0, 6)).contiguous()
        else:
            d_o = T.nchw_f32_to_nhwc_bf16(d_thin, pitch=496)
        do = self.bn_sep.bwd(d_o)
        # biases of the two 1x15 convs: both get the column sums of `do` (one gradient view, two variables)
        T.col_sums_into(do.reshape(N * fm * fm, 496), self.sep_b_bias_vec.grad)
        dmid = self.sep_b.bwd(do)
        dbackbone = self.sep_a.bwd(dmid)
        # ---- RPN head (unless the step already ran it while it waited for the proposals) ----
        if getattr(t, "d_rpn_feat", None) is None:
            self.backward_rpn_head()
        d_rpn_feat = t.d_rpn_feat
        # ---- backbone (each stage's gradient bucket goes to the communication stream as soon as it is complete) ----
        if self.xception:
            self._stage_done("heads")
            self.body.bwd(d_rpn_feat, dbackbone, stage_done=self._stage_done)
        else:
            dx = self.bn_final.bwd(dbackbone)
            self._stage_done("heads")
            for li in (3, 2, 1, 0):
                layer = self.layers[li]
                if li == 2:  # x3 also feeds the RPN feature's batch_norm_relu
                    dx = self.bn_rpn.bwd(d_rpn_feat, add_in=dx)
                for bi in range(len(layer) - 1, -1, -1):
                    blk = layer[bi]
                    dx = blk.bwd(dx)
                if li >= 2:
                    self._stage_done("layer%d" % (li + 1))
            Ho, Wo, Wimg = t.stem_geom
            dy0 = T.maxpool3x3s2_bwd(t.pool_arg, dx, t.y0_hw)
            ops.conv2d_wgrad(t.x8, dy0, 7, 7, padding=(3, 3, Ho, Wo), strides=(2, 2), cin=3, cout=64, dw=self.stem.p.dw,
                             fold_w=(Wimg, 3))
        if self.wg_stream is not None and getattr(self.wg_stream, "xdet_forked", False):
            torch.cuda.current_stream().wait_stream(self.wg_stream)   # JOIN (weight-gradient stream)
            self.wg_stream.xdet_forked = False
        if self.f32:
            self.grads.mul_(1.0 / t.S)  # exact: S is a power of two

    def apply_gradients(self, apply_update=True):
        """All-reduce (the buckets not sent yet, join) + MomentumOptimizer.apply_gradients with the L2 term folded in."""
        p = self.params
        world = self._allreduce_finish()
        if apply_update:
            lr = learning_rate(p, self.global_step)
            gs = 1.0 / world
            if self.f32:
                for c in self.convs:
                    c.update(lr, p['momentum'], p['weight_decay'], gs)
                for v in self.vecs:
                    v.update(lr, p['momentum'], p['weight_decay'], gs)
                # second 1x15 bias: same gradient as the first
                T.sgd_momentum_vec(self.sep_b_bias_vec.grad, self.sep_b_biases[1], self.sep_b_bias_mom2, lr,
                                   p['momentum'], p['weight_decay'], gs)
            else:
                # every variable in one launch (the stem's folded layout keeps its own kernel)
                if self._sgd_plan is None:
                    self._sgd_plan = self._build_sgd_plan()
                for c in self.convs:
                    if c.fold:
                        c.update(lr, p['momentum'], p['weight_decay'], gs)
                self._sgd_plan.step(lr, p['momentum'], gs)
            self.global_step += 1
        return world

    def _build_sgd_plan(self):
        p = self.params
        plan = T.SgdPlan()
        for c in self.convs:
            if not c.fold:
                for (_, t, co, ci), m in zip(c.masters, c.mom):
                    plan.add_conv(c.dw, t, m, c.pack, c.dpack, p['weight_decay'], co, ci)
        for v in self.vecs:
            for i, (t, m) in enumerate(zip(v.tensors, v.mom)):
                plan.add_vec(v.grad[i * v.seg:i * v.seg + t.numel()], t, m, p['weight_decay'] if v.decayed else 0.0)
        plan.add_vec(self.sep_b_bias_vec.grad, self.sep_b_biases[1], self.sep_b_bias_mom2, p['weight_decay'])
        return plan

    def outputs(self):
        t = self.t
        return {'rpn_cross_entropy_loss': t.rpn_ce, 'rpn_location_loss': t.rpn_loc, 'head_loss': t.head_loss,
                'rpn_idx': t.rpn_idx, 'rois_all': t.rois_all, 'roi_idx': t.roi_idx, 'ohem_idx': t.sel,
                'rois': t.rois, 'roi_labels': t.roi_lab, 'roi_targets': t.roi_tgt, 'glabels': t.glabels,
                'rpn_out': t.rpn_out, 'large_sep_feature': t.thin}


def synthetic_batch(params, batch, seed, device="cuda", max_gt=6):
    """VOC-shaped synthetic tensors: images U(-1,1); 1..6 boxes per image with side >= 0.1,
    labels 1..20; the uniform key arrays that stand in for tf.random_shuffle."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    size = params['train_image_size']
    images = torch.rand((batch, 3, size, size), generator=g) * 2 - 1
    gt = torch.zeros((batch, max_gt, 4))
    gl = torch.zeros((batch, max_gt), dtype=torch.int32)
    for n in range(batch):
        k = int(torch.randint(1, max_gt + 1, (1,), generator=g))
        cy, cx = torch.rand(k, generator=g) * 0.6 + 0.2, torch.rand(k, generator=g) * 0.6 + 0.2
        h, w = torch.rand(k, generator=g) * 0.4 + 0.1, torch.rand(k, generator=g) * 0.4 + 0.1
        gt[n, :k] = torch.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], -1).clamp(0, 1)
        gl[n, :k] = torch.randint(1, params['num_classes'], (k,), generator=g).to(torch.int32)
    fm = size // 16
    a_tot = fm * fm * 22
    R0 = params['rpn_post_nms_top_n']
    keys = {'rpn_fg': torch.rand((1, batch * a_tot), generator=g), 'rpn_bg': torch.rand((1, batch * a_tot), generator=g),
            'rpn_up': torch.rand((1, batch * params['rpn_anchors_per_image']), generator=g),
            'prop': torch.rand((batch, R0), generator=g), 'roi_fg': torch.rand((batch, R0 + max_gt), generator=g),
            'roi_bg': torch.rand((batch, R0 + max_gt), generator=g),
            'roi_up': torch.rand((batch, params['roi_one_image']), generator=g)}
    to = lambda t: t.to(device).contiguous()
    return to(images), to(gt), to(gl), {k: to(v) for k, v in keys.items()}


def arg_parser():
    """The reference's flags (light_head_rfcn_train.py:38-168), same names and defaults, as ``--name value``."""
    import argparse
    ap = argparse.ArgumentParser(description="Light-Head R-CNN training steps on synthetic VOC-shaped tensors")
    for k, v in _DEFAULTS.items():
        if isinstance(v, bool):
            ap.add_argument("--" + k, type=lambda s: s.lower() in ("1", "true", "yes"), default=v)
        elif isinstance(v, tuple):
            ap.add_argument("--" + k, type=lambda s: tuple(int(t) for t in s.split(",")), default=v)
        elif v is None:
            ap.add_argument("--" + k, type=int, default=None)
        else:
            ap.add_argument("--" + k, type=type(v), default=v)
    ap.add_argument("--steps", type=int, default=3)
    return ap


def main(argv=None):
    """The reference's ``main`` (light_head_rfcn_train.py:456-526) on synthetic VOC-shaped tensors: resume from
    ``--model_dir`` if it holds a checkpoint (what the Estimator does), else the fine-tuning restore of
    ``get_init_fn_for_scaffold`` from ``--checkpoint_path`` (scope renaming, ``--checkpoint_exclude_scopes``,
    ``--ignore_missing_vars``), else -- loudly -- random initialisation; a checkpoint is written to ``--model_dir`` every
    ``--save_checkpoints_secs`` and at the end."""
    import sys
    import time

    from .utility import train_helper as th
    args = arg_parser().parse_args(argv)
    params = make_params(**{k: getattr(args, k) for k in _DEFAULTS})
    flags = types.SimpleNamespace(**params)
    tr = LightHeadTrainer(params)
    resume = th.latest_checkpoint(flags.model_dir)
    if resume:
        print("resuming from %s" % tr.restore_checkpoint(resume))
    else:
        init_fn = th.get_init_fn_for_scaffold(flags, tr.trainable_variable_names(),
                                              shapes={k: tuple(v.shape) for k, v in tr.store.vars.items()})
        try:
            sd = init_fn() if init_fn is not None else {}
        except (FileNotFoundError, OSError) as e:
            sd = {}
            sys.stderr.write("WARNING: no checkpoint to fine-tune from (%s): TRAINING FROM RANDOM INITIALISATION\n" % e)
        for k, v in sd.items():
            tr.store.vars[k].copy_(torch.as_tensor(v).reshape(tr.store.vars[k].shape))
        if sd:
            tr.refresh_packs()
            print("restored %d variables from the fine-tuning checkpoint" % len(sd))
    batch = synthetic_batch(params, args.batch_size, seed=3)
    last_save = time.time()
    for i in range(args.steps):
        out = tr.step(*batch)
        print("step %d  rpn_ce %.4f  rpn_loc %.4f  head %.4f" % (tr.global_step, float(out['rpn_cross_entropy_loss']),
                                                                 float(out['rpn_location_loss']), float(out['head_loss'])))
        if time.time() - last_save >= flags.save_checkpoints_secs:
            print("saved %s" % tr.save_checkpoint(flags.model_dir))
            last_save = time.time()
    print("saved %s" % tr.save_checkpoint(flags.model_dir))


if __name__ == '__main__':
    main()
