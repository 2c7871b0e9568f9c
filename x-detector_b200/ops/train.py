"""Host wrappers of the training-step kernels (``csrc/train_ops.cu``): batch-norm with batch statistics and
its backward, pooling backward, losses, the momentum step, target assignment and sampling.  Torch tensors are
the container only."""
import ctypes

import torch

from .. import _native
from .conv import same_pad


def _st():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def col_stats(x2d, with_squares=True, C=None):
    """x2d [rows, cs] bf16 (or fp32: the fp32-accurate training mode) -> fp32 [2*C] (sums, sums of squares) or [C]."""
    rows, cs = x2d.shape
    C = cs if C is None else C
    sums = torch.zeros((2 if with_squares else 1) * C, dtype=torch.float32, device=x2d.device)
    col_sums_into(x2d, sums, with_squares=with_squares, C=C)
    return sums


def col_sums_into(x2d, sums, with_squares=False, C=None):
    """Column sums of x2d [rows, cs] ADDED to ``sums`` (a view of the flat gradient buffer for bias gradients)."""
    rows, cs = x2d.shape
    C = cs if C is None else C
    assert x2d.is_contiguous()
    if x2d.dtype == torch.float32:
        _native.check(_native.lib().xdet_col_stats_f32(x2d.data_ptr(), rows, C, cs, 1 if with_squares else 0, 1,
                                                       sums.data_ptr(), _st()))
    else:
        _native.check(_native.lib().xdet_col_stats_bf16(x2d.data_ptr(), rows, C, cs, 1 if with_squares else 0,
                                                        sums.data_ptr(), _st()))


def bn_relu_bwd_into(dy, x, st, relu, sums_view, add_in=None, sums_zeroed=False):
    """Backward of y = [relu](x*scale+shift) with batch statistics; ``sums_view`` ([0,cs) = dbeta, [cs,2cs) = dgamma) is
    overwritten -- usually a view of the flat gradient buffer (``sums_zeroed``: the caller vouches that it holds zeros,
    the bf16 path then skips its memset).  Returns dx (dtype of x)."""
    cs = x.shape[-1]
    assert dy.shape == x.shape and dy.dtype == x.dtype and dy.is_contiguous() and x.is_contiguous()
    assert sums_view.numel() == 2 * cs and (add_in is None or (add_in.is_contiguous() and add_in.dtype == x.dtype))
    dx = torch.empty_like(x)
    args = (dy.data_ptr(), x.data_ptr(), st.scale.data_ptr(), st.shift.data_ptr(), st.mean.data_ptr(),
            st.invstd.data_ptr(), st.rows, cs, 1 if relu else 0, _p(add_in), sums_view.data_ptr(), dx.data_ptr())
    if x.dtype == torch.float32:
        _native.check(_native.lib().xdet_bn_relu_bwd_f32(*args, _st()))
    else:
        _native.check(_native.lib().xdet_bn_relu_bwd_bf16(*args, 1 if sums_zeroed else 0, _st()))
    return dx


_BN_SCRATCH = {}


def _bn_scratch(device, C):
    """The zeroed scratch of xdet_bn_train_stats_bf16 for the current stream (the kernel hands it back zeroed; launches
    on one stream are ordered, so one buffer per stream serves every layer)."""
    key = (device.index, torch.cuda.current_stream().cuda_stream)
    buf = _BN_SCRATCH.get(key)
    need = _native.lib().xdet_bn_train_scratch_bytes(max(C, 4096))
    if buf is None or buf.numel() < need:
        buf = _BN_SCRATCH[key] = torch.zeros(need, dtype=torch.uint8, device=device)
    return buf


class BNState(object):
    """What the backward of one training-mode batch-norm needs."""
    __slots__ = ("scale", "shift", "mean", "invstd", "rows")


def bn_train(x, gamma, beta, eps, decay=None, moving_mean=None, moving_var=None):
    """Batch statistics of x [..., C] bf16 -> BNState (scale/shift normalise with the BATCH mean/variance)."""
    C = x.shape[-1]
    x2 = x.reshape(-1, C)
    st = BNState()
    st.rows = x2.shape[0]
    st.scale, st.shift, st.mean, st.invstd = torch.empty((4, C), dtype=torch.float32, device=x.device).unbind(0)
    if x.dtype == torch.bfloat16:  # statistics + finalize in one launch over a self-cleaning scratch
        assert x2.is_contiguous()
        _native.check(_native.lib().xdet_bn_train_stats_bf16(
            x2.data_ptr(), st.rows, C, C, gamma.data_ptr(), beta.data_ptr(), eps, 0.0 if decay is None else decay,
            _p(moving_mean), _p(moving_var), st.scale.data_ptr(), st.shift.data_ptr(), st.mean.data_ptr(),
            st.invstd.data_ptr(), _bn_scratch(x.device, C).data_ptr(), _st()))
        return st
    sums = col_stats(x2, True)
    _native.check(_native.lib().xdet_bn_finalize(sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), st.rows, C, eps,
                                                 0.0 if decay is None else decay, _p(moving_mean), _p(moving_var),
                                                 st.scale.data_ptr(), st.shift.data_ptr(), st.mean.data_ptr(),
                                                 st.invstd.data_ptr(), _st()))
    return st


def bn_train_apply(x, sums, gamma, beta, eps, decay=None, moving_mean=None, moving_var=None, relu=True):
    """Batch-norm (+ReLU) of x [..., C] bf16 from statistics that already exist -- ``sums`` fp32 [2*C] (sums, sums of
    squares over all rows), accumulated by the convolution that produced x (``conv2d_nhwc(stats=...)``): bn_finalize and
    the normalisation in one launch.  Returns (y, BNState)."""
    C = x.shape[-1]
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and sums.numel() == 2 * C and sums.dtype == torch.float32
    st = BNState()
    st.rows = x.numel() // C
    st.scale, st.shift, st.mean, st.invstd = torch.empty((4, C), dtype=torch.float32, device=x.device).unbind(0)
    y = torch.empty_like(x)
    _native.check(_native.lib().xdet_bn_train_apply_bf16(
        x.data_ptr(), y.data_ptr(), st.rows, C, sums.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps,
        0.0 if decay is None else decay, _p(moving_mean), _p(moving_var), st.scale.data_ptr(), st.shift.data_ptr(),
        st.mean.data_ptr(), st.invstd.data_ptr(), 1 if relu else 0, _st()))
    return y, st


def bn_relu_bwd(dy, x, st, relu=True, add_in=None):
    """-> (dx bf16 like x, dgamma [C], dbeta [C]) for y = relu(x*scale+shift) with batch statistics."""
    C = x.shape[-1]
    sums = torch.empty(2 * C, dtype=torch.float32, device=x.device)
    dx = bn_relu_bwd_into(dy, x, st, relu, sums, add_in)
    return dx, sums[C:], sums[:C]


def relu_bwd(dy, y):
    assert dy.shape == y.shape and dy.is_contiguous() and dy.dtype == y.dtype
    if getattr(y, "_pair_only", False):
        raise ValueError("relu_bwd needs the fp32 values of y (it was stored as split planes only)")
    assert y.is_contiguous()
    dx = torch.empty_like(dy)
    fn = _native.lib().xdet_relu_bwd_f32 if dy.dtype == torch.float32 else _native.lib().xdet_relu_bwd_bf16
    _native.check(fn(dy.data_ptr(), y.data_ptr(), dx.data_ptr(), dy.numel(), _st()))
    return dx


def maxpool3x3s2_fwd_train(x):
    """tf.layers.max_pooling2d(3, 2, 'SAME') on NHWC bf16 -> (pooled, argmax uint8 [N,Ho,Wo,C])."""
    N, H, W, C = x.shape
    Ho, Wo = -(-H // 2), -(-W // 2)
    out = torch.empty((N, Ho, Wo, C), dtype=x.dtype, device=x.device)
    arg = torch.empty((N, Ho, Wo, C), dtype=torch.uint8, device=x.device)
    fn = (_native.lib().xdet_maxpool3x3s2_argmax_f32 if x.dtype == torch.float32
          else _native.lib().xdet_maxpool3x3s2_argmax_bf16)
    assert x.is_contiguous()
    _native.check(fn(x.data_ptr(), out.data_ptr(), arg.data_ptr(), N, H, W, C, Ho, Wo, same_pad(H, 3, 1, 2),
                     same_pad(W, 3, 1, 2), _st()))
    return out, arg


def maxpool3x3s2_bwd(argmax, dy, in_hw):
    """dx [N,H,W,C] of the pooling above from its recorded argmax."""
    N, Ho, Wo, C = dy.shape
    H, W = in_hw
    dx = torch.empty((N, H, W, C), dtype=dy.dtype, device=dy.device)
    fn = _native.lib().xdet_maxpool3x3s2_bwd_f32 if dy.dtype == torch.float32 else _native.lib().xdet_maxpool3x3s2_bwd_bf16
    assert dy.is_contiguous()
    _native.check(fn(argmax.data_ptr(), dy.data_ptr(), dx.data_ptr(), N, H, W, C, Ho, Wo, same_pad(H, 3, 1, 2),
                     same_pad(W, 3, 1, 2), _st()))
    return dx


def nchw_f32_to_nhwc_bf16(x, pitch=None):
    """[N,C,H,W] fp32 -> [N,H,W,pitch] bf16 (channels C..pitch-1 zero)."""
    N, C, H, W = x.shape
    pitch = C if pitch is None else pitch
    out = torch.empty((N, H, W, pitch), dtype=torch.bfloat16, device=x.device)
    _native.check(_native.lib().xdet_nchw_f32_to_nhwc_bf16(x.data_ptr(), out.data_ptr(), N, C, pitch, H * W, _st()))
    return out


def affine_relu_to_nchw_f32(x, scale, shift, relu=True, C=None):
    """relu(x*scale+shift) of the first C channels of x [N,H,W,cs] bf16 -> [N,C,H,W] fp32."""
    N, H, W, cs = x.shape
    C = cs if C is None else C
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    _native.check(_native.lib().xdet_affine_relu_to_nchw_f32(x.data_ptr(), scale.data_ptr(), shift.data_ptr(),
                                                             out.data_ptr(), N, C, cs, H * W, 1 if relu else 0, _st()))
    return out


def softmax_ce(logits, labels, num_classes, row_w=None, w_all=1.0, dlogits=None):
    """logits [M, ld] fp32 (classes in the first num_classes columns), labels [M] int32 ->
    (loss_row [M], dlogits [M, dld] = w_all*row_w*(softmax-onehot), written into ``dlogits`` if given)."""
    M, ld = logits.shape
    loss = torch.empty(M, dtype=torch.float32, device=logits.device)
    if dlogits is None:
        dlogits = torch.zeros((M, ld), dtype=torch.float32, device=logits.device)
    _native.check(_native.lib().xdet_softmax_ce(logits.data_ptr(), logits.stride(0), num_classes, labels.data_ptr(),
                                                _p(row_w), w_all, M, loss.data_ptr(), dlogits.data_ptr(),
                                                dlogits.stride(0), _st()))
    return loss, dlogits


def smooth_l1(pred, target, row_w=None, w_all=1.0, dpred=None):
    """pred [M, >=4] fp32 view (row stride arbitrary), target [M,4] -> (loss_row [M], dpred)."""
    M = pred.shape[0]
    loss = torch.empty(M, dtype=torch.float32, device=pred.device)
    if dpred is None:
        dpred = torch.zeros((M, 4), dtype=torch.float32, device=pred.device)
    assert target.is_contiguous() and target.shape == (M, 4)
    _native.check(_native.lib().xdet_smooth_l1(pred.data_ptr(), pred.stride(0), target.data_ptr(), _p(row_w), w_all, M,
                                               loss.data_ptr(), dpred.data_ptr(), dpred.stride(0), _st()))
    return loss, dpred


def sgd_momentum_conv(dw, w, mom, w_pack, w_dgrad_pack, lr, momentum, wd, grad_scale=1.0, co_off=0, ci_off=0,
                      fold=False):
    """w/mom: fp32 master + accumulator in TF layout [kh,kw,cin,cout] (or [cin,units] for dense); dw: packed fp32
    gradient [Cout_total, taps, cin_pad]; w_pack / w_dgrad_pack: the bf16 packs to refresh."""
    if w.dim() == 2:
        kh = kw = 1
        cin, cout = w.shape
    else:
        kh, kw, cin, cout = w.shape
    cin_pad = dw.shape[-1]
    cout_pad = 0 if w_dgrad_pack is None else w_dgrad_pack.shape[-1] // (kh * kw)
    _native.check(_native.lib().xdet_sgd_momentum_conv(dw.data_ptr(), w.data_ptr(), mom.data_ptr(), w_pack.data_ptr(),
                                                       _p(w_dgrad_pack), cout, kh, kw, cin, co_off, ci_off, cin_pad,
                                                       cout_pad, 1 if fold else 0, lr, momentum, wd, grad_scale, _st()))


def sgd_momentum_vec(g, w, mom, lr, momentum, wd=0.0, grad_scale=1.0):
    _native.check(_native.lib().xdet_sgd_momentum_vec(g.data_ptr(), w.data_ptr(), mom.data_ptr(), w.numel(), lr,
                                                      momentum, wd, grad_scale, _st()))


class SgdItem(ctypes.Structure):
    """``xdet_sgd_item`` (include/xdet_b200.h)."""
    _fields_ = [("dw", ctypes.c_void_p), ("w", ctypes.c_void_p), ("mom", ctypes.c_void_p), ("w_pack", ctypes.c_void_p),
                ("w_dgrad_pack", ctypes.c_void_p), ("Cout", ctypes.c_int), ("taps", ctypes.c_int), ("Cin", ctypes.c_int),
                ("cin_pad", ctypes.c_int), ("cout_pad", ctypes.c_int), ("tiles_ci", ctypes.c_int),
                ("tiles_co", ctypes.c_int), ("first_block", ctypes.c_int), ("wd", ctypes.c_float),
                ("reserved", ctypes.c_int)]


class SgdPlan(object):
    """The item table of ``xdet_sgd_momentum_multi``: add the variables once, ``step`` is one launch.  The table holds raw
    device pointers: every tensor handed to ``add_*`` must stay allocated (and in place) while the plan is used."""

    def __init__(self):
        self.items, self.blocks, self.table, self.keep = [], 0, None, []

    def add_conv(self, dw, w, mom, w_pack, w_dgrad_pack, wd, co_off=0, ci_off=0):
        """Same arguments as ``sgd_momentum_conv`` (regular layout)."""
        if w.dim() == 2:
            kh = kw = 1
            cin, cout = w.shape
        else:
            kh, kw, cin, cout = w.shape
        taps, cin_pad = kh * kw, dw.shape[-1]
        cout_pad = 0 if w_dgrad_pack is None else w_dgrad_pack.shape[-1] // taps
        off = co_off * taps * cin_pad + ci_off
        it = SgdItem(dw.data_ptr() + 4 * off, w.data_ptr(), mom.data_ptr(), w_pack.data_ptr() + 2 * off,
                     None if w_dgrad_pack is None else w_dgrad_pack.data_ptr() + 2 * (ci_off * taps * cout_pad + co_off),
                     cout, taps, cin, cin_pad, cout_pad, (cin + 31) // 32, (cout + 31) // 32, self.blocks, wd, 0)
        assert w.is_contiguous() and mom.is_contiguous() and dw.is_contiguous() and w_pack.dtype == torch.bfloat16
        self.items.append(it)
        self.blocks += it.tiles_ci * it.tiles_co * taps
        self.keep += [dw, w, mom, w_pack, w_dgrad_pack]
        self.table = None

    def add_vec(self, g, w, mom, wd):
        n = w.numel()
        assert g.numel() >= n and w.is_contiguous() and mom.is_contiguous() and g.is_contiguous()
        self.items.append(SgdItem(g.data_ptr(), w.data_ptr(), mom.data_ptr(), None, None, n, 0, 0, 0, 0, 0, 0,
                                  self.blocks, wd, 0))
        self.blocks += (n + 255) // 256
        self.keep += [g, w, mom]
        self.table = None

    def step(self, lr, momentum, grad_scale=1.0):
        if not self.items:
            return
        if self.table is None:
            raw = bytes((SgdItem * len(self.items))(*self.items))
            self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.keep[0].device)
        _native.check(_native.lib().xdet_sgd_momentum_multi(self.table.data_ptr(), len(self.items), self.blocks, lr,
                                                            momentum, grad_scale, _st()))


def match_encode(boxes, gt, gt_labels, allowed_border, high_thres, low_thres, prior_scaling=(1., 1., 1., 1.),
                 ref_yxhw=None):
    """boxes [A,4] (shared by all images; pass ref_yxhw [A,4] = the anchors' centre form) or [N,A,4];
    gt [N,G,4] fp32, gt_labels [N,G] int32 (<= 0: padding) -> labels [N,A] int32, targets [N,A,4], scores [N,A]."""
    N, G = gt_labels.shape
    shared = boxes.dim() == 2
    A = boxes.shape[-2]
    dev = gt.device
    if G == 0:  # a batch without ground-truth slots: everything is background (the reference's tf.argmax would fail)
        return (torch.zeros((N, A), dtype=torch.int32, device=dev), torch.zeros((N, A, 4), dtype=torch.float32, device=dev),
                torch.zeros((N, A), dtype=torch.float32, device=dev))
    labels = torch.empty((N, A), dtype=torch.int32, device=dev)
    targets = torch.empty((N, A, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((N, A), dtype=torch.float32, device=dev)
    lib = _native.lib()
    ws = torch.empty(max(8, lib.xdet_match_workspace_bytes(N, G)), dtype=torch.uint8, device=dev)
    ps = (ctypes.c_float * 4)(*prior_scaling)
    assert boxes.is_contiguous() and gt.is_contiguous() and gt_labels.dtype == torch.int32 and gt_labels.is_contiguous()
    _native.check(lib.xdet_match_encode(boxes.data_ptr(), 0 if shared else A * 4, _p(ref_yxhw), gt.data_ptr(),
                                        gt_labels.data_ptr(), N, A, G, allowed_border, high_thres, low_thres, ps,
                                        labels.data_ptr(), targets.data_ptr(), scores.data_ptr(), ws.data_ptr(), _st()))
    return labels, targets, scores


def sample_fg_bg(labels, scores, bg_low, exp_fg, total, keys_fg, keys_bg, keys_up):
    """labels [groups, n] int32 (scores [groups, n] or None) -> (indices [groups, total] int32, counts [groups, 3])."""
    groups, n = labels.shape
    dev = labels.device
    if total > 4096:
        raise ValueError("sample_fg_bg draws at most 4096 samples per group (%d asked): the RPN sampler runs over the "
                         "flattened per-rank batch, batch * rpn_anchors_per_image <= 4096 (16 images at the reference's "
                         "256); shard larger batches over more ranks" % total)
    out = torch.empty((groups, total), dtype=torch.int32, device=dev)
    counts = torch.empty((groups, 3), dtype=torch.int32, device=dev)
    ws = torch.empty((groups, 2 * n), dtype=torch.int32, device=dev)
    for k in (keys_fg, keys_bg):
        assert k.shape == (groups, n) and k.dtype == torch.float32 and k.is_contiguous()
    assert keys_up.shape == (groups, total) and keys_up.is_contiguous()
    _native.check(_native.lib().xdet_sample_fg_bg(labels.data_ptr(), _p(scores), bg_low, groups, n, exp_fg, total,
                                                  keys_fg.data_ptr(), keys_bg.data_ptr(), keys_up.data_ptr(),
                                                  ws.data_ptr(), out.data_ptr(), counts.data_ptr(), _st()))
    return out, counts
