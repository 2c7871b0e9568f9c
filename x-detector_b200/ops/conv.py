"""Host side of the tensor-core convolution / GEMM (``csrc/conv_gemm.cu`` through
``xdet_conv2d_bf16``).  Tensors are NHWC bf16 CUDA torch tensors (container only).

TensorFlow layers this stands in for on the Light-Head R-CNN path: ``tf.layers.conv2d`` with
``padding='SAME'|'VALID'``, ``strides=1`` and optional ``dilation_rate`` (net/resnet_v2.py:89-100,
net/xdet_body.py:28-37, net/xception_body.py:381-400,450-475), and ``tf.layers.dense``
(net/xception_body.py:540-558).  Batch-norm (inference form) and ReLU are folded into the epilogue.
"""
import ctypes

import torch

from .. import _native


class ConvDesc(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("Cin", ctypes.c_int), ("in_cs", ctypes.c_int),
        ("Cout", ctypes.c_int), ("KH", ctypes.c_int), ("KW", ctypes.c_int), ("dil_h", ctypes.c_int),
        ("dil_w", ctypes.c_int), ("pad_top", ctypes.c_int), ("pad_left", ctypes.c_int),
        ("Hout", ctypes.c_int), ("Wout", ctypes.c_int),
        ("stride_h", ctypes.c_int), ("stride_w", ctypes.c_int), ("fold_w", ctypes.c_int), ("in_wp", ctypes.c_int),
        ("weights", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("relu", ctypes.c_int),
        ("residual", ctypes.c_void_p), ("out", ctypes.c_void_p), ("out_fp32", ctypes.c_int),
        ("out_sn", ctypes.c_longlong), ("out_sy", ctypes.c_longlong), ("out_sx", ctypes.c_longlong),
        ("out_sc", ctypes.c_longlong),
        ("out2", ctypes.c_void_p), ("scale2", ctypes.c_void_p), ("bias2", ctypes.c_void_p), ("block_n", ctypes.c_int),
        ("epi_groups", ctypes.c_int), ("max_ctas", ctypes.c_int), ("stats", ctypes.c_void_p),
    ]


# Arithmetic of the convolutions / dense layers:
#   "bf16"   (default, the throughput path) bf16 operands, fp32 accumulate;
#   "f16x2"  fp32-ACCURATE mode (csrc/conv_gemm_f16x2.cu), the precision the parity claim is benchmarked at:
#            activations stay fp32 between layers; every operand is carried as two fp16 values (hi, lo*2^11), one
#            kernel launch per layer issues three tensor-core products into two TMEM accumulators and flushes them
#            into fp32 registers with round-to-nearest every <= 768 reduction elements -> fp32-level error.
#   "fp32x3" the first parity mode (csrc/parity_ops.cu): three bf16 pieces per operand, one convolution over 6*Cin
#            channels through the bf16 kernel, long reductions as several launches.  Kept as a second opinion
#            for the tests; ~20x slower than bf16.
PRECISION = "bf16"
PRECISIONS = ("bf16", "f16x2", "fp32x3")
_chunk_cache = {}


class precision(object):
    """``with precision("fp32x3"): ...`` -- weights packed and convolutions issued inside use that arithmetic.
    A VariableStore caches packed weights, so use one store (one model object) per precision."""

    def __init__(self, mode):
        assert mode in PRECISIONS, mode
        self.mode = mode

    def __enter__(self):
        global PRECISION
        self.prev, PRECISION = PRECISION, self.mode

    def __exit__(self, *exc):
        global PRECISION
        PRECISION = self.prev


def split3_values(w):
    """fp32 tensor -> (hi, mid, lo): bf16-representable fp32 tensors with hi + mid + lo == w up to 2^-24 |w|."""
    w = w.float()
    hi = w.to(torch.bfloat16).float()
    r = w - hi  # exact
    mid = r.to(torch.bfloat16).float()
    lo = (r - mid).to(torch.bfloat16).float()
    return hi, mid, lo


class PairWeight(object):
    """Weights of one convolution / dense layer in "f16x2" precision: fp16 planes [2][Cout][taps*ceil(Cin/64)*64]
    (hi, lo*2^11) of ``w * 2^e[c]`` -- a per-output-channel power of two that lifts the channel's largest weight into
    [2^8, 2^9), far from fp16's subnormals -- and ``winv`` = 2^-e[c], which the epilogue's scale vector absorbs
    (exact).  ``fold_cs``: the fold_w layout [Cout][KH][64] of the few-channel stem."""

    def __init__(self, w_oihw, fold_cs=None):
        w = w_oihw.float()
        cout, cin, kh, kw = w.shape
        amax = w.abs().reshape(cout, -1).amax(dim=1).clamp_min(1e-30)
        e = torch.floor(8.0 - torch.log2(amax))
        sc = torch.exp2(e)
        self.winv = torch.exp2(-e).contiguous()
        w = w * sc.view(-1, 1, 1, 1)  # exact (power of two)
        if fold_cs is None:
            cpad = (cin + 63) // 64 * 64
            flat = torch.zeros((cout, kh * kw, cpad), dtype=torch.float32, device=w.device)
            flat[:, :, :cin] = w.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin)
            flat = flat.reshape(cout, kh * kw * cpad)
        else:
            assert cin <= fold_cs and kw * fold_cs <= 64
            flat = torch.zeros((cout, kh, 64), dtype=torch.float32, device=w.device)
            flat[:, :, :kw * fold_cs].view(cout, kh, kw, fold_cs)[..., :cin] = w.permute(0, 2, 3, 1)
            flat = flat.reshape(cout, kh * 64)
        hi = flat.to(torch.float16)
        lo = ((flat - hi.float()) * 2048.0).to(torch.float16)
        self.planes = torch.stack([hi, lo]).contiguous()
        self.cout = cout
        self._scales = {}

    def data_ptr(self):
        return self.planes.data_ptr()

    def scale_eff(self, scale):
        """``scale * winv`` (exact), cached per scale tensor (recomputed if the tensor was written since)."""
        if scale is None:
            return self.winv
        key = scale.data_ptr()
        hit = self._scales.get(key)
        if hit is None or hit[0] is not scale or hit[1] != scale._version:
            hit = (scale, scale._version, (scale.float() * self.winv).contiguous())
            self._scales[key] = hit
        return hit[2]


class ConvF16x2Desc(ctypes.Structure):
    _fields_ = [
        ("N", ctypes.c_int), ("H", ctypes.c_int), ("W", ctypes.c_int), ("Cin", ctypes.c_int), ("in_cs", ctypes.c_int),
        ("Cout", ctypes.c_int), ("KH", ctypes.c_int), ("KW", ctypes.c_int), ("dil_h", ctypes.c_int),
        ("dil_w", ctypes.c_int), ("pad_top", ctypes.c_int), ("pad_left", ctypes.c_int),
        ("Hout", ctypes.c_int), ("Wout", ctypes.c_int),
        ("stride_h", ctypes.c_int), ("stride_w", ctypes.c_int), ("fold_w", ctypes.c_int), ("in_wp", ctypes.c_int),
        ("in_plane", ctypes.c_longlong), ("weights", ctypes.c_void_p), ("w_plane", ctypes.c_longlong),
        ("scale", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("relu", ctypes.c_int),
        ("residual", ctypes.c_void_p), ("out", ctypes.c_void_p),
        ("out_sn", ctypes.c_longlong), ("out_sy", ctypes.c_longlong), ("out_sx", ctypes.c_longlong),
        ("out_sc", ctypes.c_longlong),
        ("out_pair", ctypes.c_void_p), ("pair_plane", ctypes.c_longlong), ("pair_cs", ctypes.c_int),
        ("out2", ctypes.c_void_p), ("scale2", ctypes.c_void_p), ("bias2", ctypes.c_void_p),
        ("out2_pair", ctypes.c_void_p),
        ("block_n", ctypes.c_int), ("chunk_kb", ctypes.c_int), ("max_ctas", ctypes.c_int), ("cluster", ctypes.c_int),
    ]


# 2-CTA clusters that share the B tile (xdet_conv_f16x2_desc.cluster): 1 = never (default), 2 = always, 0 = the tuner
# decides.  Measured on the benchmarked step: bit-identical results, a quarter less operand traffic, and
# no gain -- the mid-size layers are bound by latency x bytes in flight (3 stages of 64 KB), not by L2 bandwidth --
# so it stays off and out of the tuner's candidate list.
F16X2_CLUSTER = 1
# k-blocks (64 reduction elements each) per accumulator flush of the f16x2 kernel; 0 = the library default (2)
F16X2_CHUNK_KB = 0
# the f16x2 epilogue also writes the split planes of its NHWC outputs (the next convolution then needs no split pass)
F16X2_FUSE_SPLIT = True


def split2(x, cin=None, relu=False, pad_w=None):
    """fp32 [N,H,W,C] (any strides) -> f16x2 planes [2,N,H,Wp,ceil8(C)] fp16: hi = fp16(v), lo = fp16((v-hi)*2^11)
    (csrc/conv_gemm_f16x2.cu).  ``pad_w`` = (Wp, x_off): row-padded layout for the fold_w stem (zero padding)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
    N, H, W, C = x.shape
    cin = C if cin is None else cin
    cs = (cin + 7) // 8 * 8
    if pad_w is None:
        Wp, x_off = W, 0
        out = torch.empty((2, N, H, Wp, cs), dtype=torch.float16, device=x.device)
    else:
        # the padding is zero-filled ONCE: the buffer is cached per shape and only its interior is rewritten
        Wp, x_off = pad_w
        key = ("padded", N, H, Wp, cs, x_off, W, str(x.device))
        out = _chunk_cache.get(key)
        if out is None:
            out = torch.zeros((2, N, H, Wp, cs), dtype=torch.float16, device=x.device)
            _chunk_cache[key] = out
    sn, sy, sx, sc = x.stride()
    with torch.cuda.device(x.device):
        rc = _native.lib().xdet_split2_f16(x.data_ptr(), sn, sy, sx, sc, N, H, W, cin, out.data_ptr(), cs, Wp, x_off,
                                           out.stride(0), 1 if relu else 0, torch.cuda.current_stream().cuda_stream)
    _native.check(rc)
    return out


def pair_of(x, cin=None):
    """The f16x2 planes of an fp32 NHWC activation: the ones its producer attached, else a split pass."""
    p = getattr(x, "_pair", None)
    if p is not None and p.shape[1:4] == x.shape[0:3]:
        return p
    p = split2(x, cin)
    try:
        x._pair = p  # a second consumer of the same tensor (projection shortcut + first conv) reuses it
    except AttributeError:
        pass
    return p


def pair_only(shape, planes):
    """Stand-in for an activation that exists only as f16x2 planes (every consumer is a convolution): right shape and
    dtype for the builders' bookkeeping, no fp32 storage behind it."""
    t = torch.empty((1,), dtype=torch.float32, device=planes.device).expand(shape)
    t._pair = planes
    t._pair_only = True
    return t



def pack_conv_weight(w_oihw):
    """[Cout, Cin, KH, KW] float -> bf16 [Cout, KH*KW*ceil(Cin/64)*64] (tap-major, channels zero-padded).
    In "fp32x3" precision the input-channel axis becomes the six blocks [mid|hi|lo|hi|mid|hi] that pair with the
    activation blocks [mid|lo|hi|mid|hi|hi] written by ``split3``."""
    if PRECISION == "f16x2":
        return PairWeight(w_oihw)
    if PRECISION == "fp32x3":
        hi, mid, lo = split3_values(w_oihw)
        w_oihw = torch.cat([mid, hi, lo, hi, mid, hi], dim=1)
    cout, cin, kh, kw = w_oihw.shape
    cpad = (cin + 63) // 64 * 64
    w = torch.zeros((cout, kh * kw, cpad), dtype=torch.float32, device=w_oihw.device)
    w[:, :, :cin] = w_oihw.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin).float()
    return w.reshape(cout, kh * kw * cpad).to(torch.bfloat16).contiguous()


def pack_fold_weight(w_oihw, cs=8):
    """[Cout, Cin<=cs, KH, KW] float -> bf16 [Cout, KH*64] for the fold_w mode: element kw*cs + ci of filter row kh.
    ("fp32x3" precision has no fold mode: the generic split pack is returned and ``conv2d_image_fold`` runs the
    generic strided convolution.)"""
    if PRECISION == "f16x2":
        return PairWeight(w_oihw, fold_cs=cs)
    if PRECISION == "fp32x3":
        return pack_conv_weight(w_oihw)
    cout, cin, kh, kw = w_oihw.shape
    assert cin <= cs and kw * cs <= 64
    w = torch.zeros((cout, kh, 64), dtype=torch.float32, device=w_oihw.device)
    w[:, :, :kw * cs].view(cout, kh, kw, cs)[..., :cin] = w_oihw.permute(0, 2, 3, 1).float()
    return w.reshape(cout, kh * 64).to(torch.bfloat16).contiguous()


def same_pad(n, k, dil=1, stride=1):
    """TF 'SAME' low-side padding for extent n: total = max((ceil(n/s)-1)*s + k_eff - n, 0), low = total//2."""
    k_eff = (k - 1) * dil + 1
    out = -(-n // stride)
    total = max((out - 1) * stride + k_eff - n, 0)
    return total // 2


def _ptr(t):
    return None if t is None else t.data_ptr()


# When set to a list, every launch appends (start_event, end_event, algorithmic_flops, shape tuple):
# bench.py uses it to time the conv/GEMM kernel live and to count algorithmic FLOPs (2*MACs).
PROFILE = None

# Upper bound on the persistent CTAs of every convolution launched while set (0 = one per SM): the model_fn
# lowers it while the proposal stream runs beside the backbone so that its one-CTA-per-image kernels find free SMs.
MAX_CTAS = 0


# One-time tile-shape tuning per distinct layer shape (like cudnn.benchmark): with AUTOTUNE set, the first eager call
# of a shape times the N-tile widths {heuristic, 64, 128, 256} x {1, 2} epilogue warpgroups (10 back-to-back launches
# replayed from a CUDA graph each) and later calls -- including CUDA-graph captures -- use the fastest.  The choice
# does not change results: every output element accumulates its K products in the same order for any tile width.
AUTOTUNE = False
_tune_cache = {}


def _autotune(x, d, reps=10):
    lib = _native.lib()
    cands = [(0, 0)] + [(bn, eg) for bn in (64, 128, 256) for eg in (1, 2) if bn <= max(64, (d.Cout + 63) // 64 * 64)]
    best, best_t = (0, 0), None
    cur = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=x.device)
    for bn, eg in cands:
        d.block_n, d.epi_groups = bn, eg
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            if lib.xdet_conv2d_bf16(x.data_ptr(), ctypes.byref(d), side.cuda_stream) != 0:
                continue  # this tile shape is not available for the layer (shared memory, alignment)
            side.synchronize()
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for _ in range(reps):
                        lib.xdet_conv2d_bf16(x.data_ptr(), ctypes.byref(d), torch.cuda.current_stream().cuda_stream)
                g.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(side)
                g.replay()
                e1.record(side)
                side.synchronize()
                t = e0.elapsed_time(e1)
            except RuntimeError:
                continue
        if best_t is None or t < best_t * 0.985:  # prefer earlier candidates (the heuristic) on ties
            best, best_t = (bn, eg), t
    cur.wait_stream(side)
    d.block_n, d.epi_groups = 0, 0
    return best


def _autotune_f16x2(planes, d, reps=10):
    """Tile width and cluster mode of the f16x2 kernel for one layer shape: the heuristic's choice against N tiles of
    32 / 64 / 128, independent CTAs against 2-CTA clusters that share the B tile, each timed as ``reps`` launches
    replayed from a CUDA graph (results do not depend on the choice: same products, same accumulation order)."""
    lib = _native.lib()
    best, best_t = (0, 1), None
    cur = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=planes.device)
    fixed = d.cluster
    cands = [(bn, cl) for cl in ((1, 2) if fixed == 0 else (fixed,)) for bn in (0, 128, 64, 32)]
    for bn, cl in cands:
        d.block_n, d.cluster = bn, cl
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            if lib.xdet_conv2d_f16x2(planes.data_ptr(), ctypes.byref(d), side.cuda_stream) != 0:
                continue
            side.synchronize()
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    for _ in range(reps):
                        lib.xdet_conv2d_f16x2(planes.data_ptr(), ctypes.byref(d), torch.cuda.current_stream().cuda_stream)
                g.replay()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(side)
                g.replay()
                e1.record(side)
                side.synchronize()
                t = e0.elapsed_time(e1)
            except RuntimeError:
                continue
        if best_t is None or t < best_t * 0.985:  # prefer earlier candidates (the heuristic) on ties
            best, best_t = (bn, cl), t
    cur.wait_stream(side)
    d.block_n, d.cluster = 0, fixed
    return best


class cta_limit(object):
    def __init__(self, n):
        self.n = n

    def __enter__(self):
        global MAX_CTAS
        self.prev, MAX_CTAS = MAX_CTAS, self.n

    def __exit__(self, *exc):
        global MAX_CTAS
        MAX_CTAS = self.prev


def conv2d_nhwc(x, w_packed, cout, kh, kw, *, dilation=(1, 1), padding="SAME", scale=None, bias=None, relu=False,
                residual=None, out=None, out_layout="nhwc_bf16", out2=None, scale2=None, bias2=None, cin=None,
                block_n=0, strides=(1, 1), fold_w=None, skip_out=False, epi_groups=0, forms="both", forms2="both",
                stats=None):
    """x: [N,H,W,C] bf16 (channel stride may be padded: pass the true ``cin``).  Returns the output tensor
    ([N,Ho,Wo,Cout] bf16 for 'nhwc_bf16', [N,Cout,Ho,Wo] fp32 for 'nchw_f32', [N,Ho,Wo,Cout] fp32 for 'nhwc_f32').
    ``forms`` / ``forms2`` (read by the "f16x2" precision only): which forms of ``out`` / ``out2`` their consumers need
    -- "both", "f32" (no convolution reads it: no split planes) or "pair" (only convolutions read it: no fp32 copy).
    ``stats`` (bf16 precision, bf16 NHWC output): fp32 [2*cout] the kernel ADDS the per-channel sums and sums of squares of
    the stored output to -- the batch statistics of the batch-norm that reads it (ops.train.bn_train_apply)."""
    if stats is not None and (isinstance(w_packed, PairWeight) or x.dtype != torch.bfloat16):
        raise ValueError("conv2d_nhwc(stats=...) is a bf16-precision feature")
    if isinstance(w_packed, PairWeight):
        return _conv2d_f16x2(x, w_packed, cout, kh, kw, dilation=dilation, padding=padding, scale=scale, bias=bias,
                             relu=relu, residual=residual, out=out, out_layout=out_layout, out2=out2, scale2=scale2,
                             bias2=bias2, cin=cin, block_n=block_n, strides=strides, fold_w=fold_w, skip_out=skip_out,
                             forms=forms, forms2=forms2)
    if x.dtype == torch.float32:  # parity mode: fp32 activations
        return _conv2d_fp32x3(x, w_packed, cout, kh, kw, dilation=dilation, padding=padding, scale=scale, bias=bias,
                              relu=relu, residual=residual, out=out, out_layout=out_layout, out2=out2, scale2=scale2,
                              bias2=bias2, cin=cin, block_n=block_n, strides=strides, skip_out=skip_out)
    assert x.is_cuda and x.dtype == torch.bfloat16 and x.dim() == 4
    N, H, W, cs = x.shape
    assert x.is_contiguous()
    cin = cs if cin is None else cin
    dh, dw = dilation
    sh, sw = strides
    in_wp = 0
    if fold_w is not None:
        # x is the horizontally padded few-channel image [N,H,in_wp,cs]; fold_w = (true width, pad_left)
        in_wp = W
        W = fold_w[0]
    if padding == "SAME":
        pt, pl = same_pad(H, kh, dh, sh), same_pad(W, kw, dw, sw)
        Ho, Wo = -(-H // sh), -(-W // sw)
    elif padding == "VALID":
        pt = pl = 0
        Ho, Wo = (H - (kh - 1) * dh - 1) // sh + 1, (W - (kw - 1) * dw - 1) // sw + 1
    else:
        pt, pl, Ho, Wo = padding  # explicit (pad_top, pad_left, Hout, Wout)
    if fold_w is not None:
        assert pl == fold_w[1], "the materialised left padding must equal the convolution's"
    dev = x.device
    if skip_out:
        assert out2 is not None and out is None and out_layout == "nhwc_bf16"
        out = out2  # geometry only; the kernel stores nothing through `out`
    if out is None:
        if out_layout == "nhwc_bf16":
            out = torch.empty((N, Ho, Wo, cout), dtype=torch.bfloat16, device=dev)
        elif out_layout == "nhwc_f32":
            out = torch.empty((N, Ho, Wo, cout), dtype=torch.float32, device=dev)
        elif out_layout == "nchw_f32":
            out = torch.empty((N, cout, Ho, Wo), dtype=torch.float32, device=dev)
        else:
            raise ValueError(out_layout)
    if out_layout == "nchw_f32":
        sn, sc, sy, sx = out.stride()
    else:
        sn, sy, sx, sc = out.stride()
    d = ConvDesc(N, H, W, cin, cs, cout, kh, kw, dh, dw, pt, pl, Ho, Wo, sh, sw, 0 if fold_w is None else 1, in_wp,
                 w_packed.data_ptr(), _ptr(scale), _ptr(bias),
                 1 if relu else 0, _ptr(residual), None if skip_out else out.data_ptr(),
                 0 if out.dtype == torch.bfloat16 else 1, sn, sy, sx, sc, _ptr(out2), _ptr(scale2), _ptr(bias2),
                 block_n, epi_groups, MAX_CTAS, _ptr(stats))
    if stats is not None:
        assert stats.dtype == torch.float32 and stats.numel() == 2 * cout and stats.is_contiguous()
    with torch.cuda.device(dev):
        if AUTOTUNE and block_n == 0 and epi_groups == 0:
            key = (N, H, W, cin, cs, cout, kh, kw, dh, dw, pt, pl, Ho, Wo, sh, sw, fold_w is not None, bool(relu),
                   residual is not None, out2 is not None, bool(skip_out), out_layout, MAX_CTAS > 0, stats is not None)
            if key not in _tune_cache and not torch.cuda.is_current_stream_capturing():
                if stats is not None:  # (the timing runs add to a scratch copy, not to the caller's statistics)
                    scratch = torch.zeros_like(stats)
                    d.stats = scratch.data_ptr()
                _tune_cache[key] = _autotune(x, d)
                d.stats = _ptr(stats)
            d.block_n, d.epi_groups = _tune_cache.get(key, (0, 0))
        if PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = _native.lib().xdet_conv2d_bf16(x.data_ptr(), ctypes.byref(d), torch.cuda.current_stream().cuda_stream)
        if PROFILE is not None:
            e1.record()
            PROFILE.append((e0, e1, 2.0 * N * Ho * Wo * cout * cin * kh * kw, (N, H, W, cin, cout, kh, kw)))
    _native.check(rc)
    return None if skip_out else out


def split3(x, cin=None):
    """fp32 [N,H,W,C] (any strides) -> bf16 [N,H,W,ceil8(6*C)]: blocks [mid|lo|hi|mid|hi|hi] of the three bf16 pieces of
    every value (csrc/parity_ops.cu)."""
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 4
    N, H, W, C = x.shape
    cin = C if cin is None else cin
    out_cs = (6 * cin + 7) // 8 * 8
    out = torch.empty((N, H, W, out_cs), dtype=torch.bfloat16, device=x.device)
    sn, sy, sx, sc = x.stride()
    with torch.cuda.device(x.device):
        rc = _native.lib().xdet_split3_bf16(x.data_ptr(), sn, sy, sx, sc, N, H, W, cin, out.data_ptr(), out_cs,
                                            torch.cuda.current_stream().cuda_stream)
    _native.check(rc)
    return out


def f32_post(x, residual=None, relu=False, out=None, scale2=None, bias2=None, out2=None, relu2=True, store_out=True):
    """fp32 elementwise tail of a parity-mode layer: v = x (+ residual) (ReLU) -> ``out`` (allocated unless
    ``store_out`` is False); ``out2`` = ReLU(v*scale2 + bias2) when scale2 is given.  Returns (out, out2)."""
    assert x.dtype == torch.float32 and x.is_contiguous()
    C = x.shape[-1]
    if store_out and out is None:
        out = torch.empty_like(x)
    if scale2 is not None and out2 is None:
        out2 = torch.empty_like(x)
    if residual is not None:
        assert residual.dtype == torch.float32 and residual.is_contiguous() and residual.shape == x.shape
    with torch.cuda.device(x.device):
        rc = _native.lib().xdet_f32_post(x.data_ptr(), _ptr(residual), 1 if relu else 0, _ptr(out) if store_out else None,
                                         _ptr(scale2), _ptr(bias2), 1 if relu2 else 0, _ptr(out2), x.numel() // C, C,
                                         torch.cuda.current_stream().cuda_stream)
    _native.check(rc)
    return (out if store_out else None), out2


# The tensor core's fp32 accumulator truncates instead of rounding, so the error of one launch grows with the number
# of accumulation steps (K/16).  Parity mode therefore caps the reduction length of a launch at PARITY_MAX_K operand
# elements (6 * taps * channels): longer reductions run as several launches over channel chunks whose fp32 partial
# results are added with round-to-nearest by xdet_f32_post.  0 = never chunk.
PARITY_MAX_K = 768


def _weight_chunk(w_packed, cout, taps, cin, c0, c1):
    """Packed "fp32x3" weights of input channels [c0, c1): the six blocks of the chunk, re-padded to 64 per tap."""
    key = (w_packed.data_ptr(), c0, c1)
    hit = _chunk_cache.get(key)
    if hit is None or hit[0] is not w_packed:
        cpad = w_packed.shape[1] // taps
        idx = torch.cat([torch.arange(j * cin + c0, j * cin + c1, device=w_packed.device) for j in range(6)])
        wv = w_packed.view(cout, taps, cpad)[:, :, idx]
        n = wv.shape[2]
        npad = (n + 63) // 64 * 64
        out = torch.zeros((cout, taps, npad), dtype=w_packed.dtype, device=w_packed.device)
        out[:, :, :n] = wv
        hit = (w_packed, out.reshape(cout, taps * npad).contiguous())
        _chunk_cache[key] = hit
    return hit[1]


def _ones(n, device):
    key = ("ones", n, str(device))
    if key not in _chunk_cache:
        _chunk_cache[key] = torch.ones(n, dtype=torch.float32, device=device)
    return _chunk_cache[key]


def _conv2d_fp32x3(x, w_packed, cout, kh, kw, *, dilation, padding, scale, bias, relu, residual, out, out_layout, out2,
                   scale2, bias2, cin, block_n, strides, skip_out):
    """``conv2d_nhwc`` in "fp32x3" precision: split the fp32 input, run the bf16 kernel over 6*Cin channels with an
    fp32 output, then (residual / second output) the fp32 elementwise tail.  Outputs are fp32."""
    cin = x.shape[-1] if cin is None else cin
    layout = "nhwc_f32" if out_layout == "nhwc_bf16" else out_layout
    tail = residual is not None or out2 is not None or scale2 is not None
    assert not tail or layout == "nhwc_f32"
    taps = kh * kw
    c_chunk = cin if PARITY_MAX_K <= 0 else max(8, PARITY_MAX_K // (6 * taps) // 8 * 8)
    if cin <= c_chunk:
        raw = conv2d_nhwc(split3(x, cin), w_packed, cout, kh, kw, dilation=dilation, padding=padding, scale=scale,
                          bias=bias, relu=relu and residual is None, out=None if tail else out, out_layout=layout,
                          cin=6 * cin, block_n=block_n, strides=strides)
    else:
        acc = None
        for c0 in range(0, cin, c_chunk):
            c1 = min(cin, c0 + c_chunk)
            part = conv2d_nhwc(split3(x[..., c0:c1]), _weight_chunk(w_packed, cout, taps, cin, c0, c1), cout, kh, kw,
                               dilation=dilation, padding=padding, out_layout="nhwc_f32", cin=6 * (c1 - c0),
                               block_n=block_n, strides=strides)
            acc = part if acc is None else f32_post(acc, residual=part, out=acc)[0]
        if scale is not None or bias is not None or relu:
            sc = scale if scale is not None else _ones(cout, x.device)
            bi = bias if bias is not None else torch.zeros_like(sc)
            raw = f32_post(acc, scale2=sc, bias2=bi, relu2=relu and residual is None, out2=acc, store_out=False)[1]
        else:
            raw = acc
        if layout == "nchw_f32":  # layout plumbing of the chunked form only (the single launch stores NCHW itself)
            raw = raw.permute(0, 3, 1, 2).contiguous()
            if out is not None:
                out.copy_(raw)
                raw = out
        elif out is not None and not tail:
            out.copy_(raw)
            raw = out
    if not tail:
        return raw
    y, _ = f32_post(raw, residual=residual, relu=relu, out=raw if out is None else out, scale2=scale2, bias2=bias2,
                    out2=out2, store_out=not skip_out)
    return y


def _conv2d_f16x2(x, w, cout, kh, kw, *, dilation, padding, scale, bias, relu, residual, out, out_layout, out2, scale2,
                  bias2, cin, block_n, strides, fold_w, skip_out, forms="both", forms2="both"):
    """``conv2d_nhwc`` in "f16x2" precision.  ``x``: fp32 NHWC activation (its split planes are taken from the
    producer when attached, else made here), or -- fold_w mode -- the row-padded f16x2 image planes themselves.
    Outputs are fp32 (the NHWC ones carry their own split planes for the next convolution); see ``forms``.
    ``out2``: a caller-provided fp32 buffer, or True = allocate what ``forms2`` asks for (then (out, out2) is returned)."""
    dh, dw_ = dilation
    sh, sw = strides
    if fold_w is not None:
        planes = x  # [2,N,H,in_wp,8]
        _, N, H, in_wp, cs = planes.shape
        W = fold_w[0]
        cin = cs if cin is None else cin
    else:
        assert x.dtype == torch.float32 and x.dim() == 4
        N, H, W, C = x.shape
        cin = C if cin is None else cin
        planes = pair_of(x, cin)
        cs = planes.shape[-1]
        in_wp = 0
    if padding == "SAME":
        pt, pl = same_pad(H, kh, dh, sh), same_pad(W, kw, dw_, sw)
        Ho, Wo = -(-H // sh), -(-W // sw)
    elif padding == "VALID":
        pt = pl = 0
        Ho, Wo = (H - (kh - 1) * dh - 1) // sh + 1, (W - (kw - 1) * dw_ - 1) // sw + 1
    else:
        pt, pl, Ho, Wo = padding
    if fold_w is not None:
        assert pl == fold_w[1], "the materialised left padding must equal the convolution's"
    dev = planes.device
    nhwc = out_layout != "nchw_f32"
    if not (F16X2_FUSE_SPLIT and nhwc):
        forms = forms2 = "f32"
    shape = (N, Ho, Wo, cout) if nhwc else (N, cout, Ho, Wo)
    want2 = out2 is not None
    alloc2 = out2 is True
    if alloc2:
        out2 = None
    if skip_out:
        assert want2 and out is None
    # fp32 outputs (caller-provided buffers are always stored in fp32)
    f32_out = f32_out2 = None
    if not skip_out and (out is not None or forms != "pair"):
        f32_out = out if out is not None else torch.empty(shape, dtype=torch.float32, device=dev)
    if want2 and (out2 is not None or forms2 != "pair"):
        f32_out2 = out2 if out2 is not None else torch.empty(shape, dtype=torch.float32, device=dev)
    geo = f32_out if f32_out is not None else (f32_out2 if f32_out2 is not None else residual)
    if geo is not None and not getattr(geo, "_pair_only", False):
        st = geo.stride()
    else:
        st = (Ho * Wo * cout, Wo * cout, cout, 1)
    if nhwc:
        sn, sy, sx, sc = st
    else:
        sn, sc, sy, sx = st
    for t in (residual, f32_out, f32_out2):
        assert t is None or (t.dtype == torch.float32 and tuple(t.shape) == shape and t.stride() == st and
                             not getattr(t, "_pair_only", False))
    pcs = (cout + 7) // 8 * 8
    pair = pair2 = None
    if not skip_out and forms != "f32":
        pair = torch.empty((2, N, Ho, Wo, pcs), dtype=torch.float16, device=dev)
    if want2 and forms2 != "f32":
        pair2 = torch.empty((2, N, Ho, Wo, pcs), dtype=torch.float16, device=dev)
    some_pair = pair if pair is not None else pair2
    d = ConvF16x2Desc(N, H, W, cin, cs, cout, kh, kw, dh, dw_, pt, pl, Ho, Wo, sh, sw, 0 if fold_w is None else 1, in_wp,
                      planes.stride(0), w.planes.data_ptr(), w.planes.stride(0), w.scale_eff(scale).data_ptr(),
                      _ptr(bias), 1 if relu else 0, _ptr(residual), _ptr(f32_out), sn, sy, sx, sc, _ptr(pair),
                      0 if some_pair is None else some_pair.stride(0), pcs, _ptr(f32_out2), _ptr(scale2), _ptr(bias2),
                      _ptr(pair2), block_n, F16X2_CHUNK_KB, MAX_CTAS, F16X2_CLUSTER)
    with torch.cuda.device(dev):
        if AUTOTUNE and block_n == 0:
            key = ("f16x2", N, H, W, cin, cs, cout, kh, kw, dh, dw_, pt, pl, Ho, Wo, sh, sw, fold_w is not None,
                   residual is not None, f32_out is not None, pair is not None, f32_out2 is not None, pair2 is not None,
                   nhwc, MAX_CTAS > 0)
            if key not in _tune_cache and not torch.cuda.is_current_stream_capturing():
                _tune_cache[key] = _autotune_f16x2(planes, d)
            d.block_n, d.cluster = _tune_cache.get(key, (0, F16X2_CLUSTER))
        if PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = _native.lib().xdet_conv2d_f16x2(planes.data_ptr(), ctypes.byref(d), torch.cuda.current_stream().cuda_stream)
        if PROFILE is not None:
            e1.record()
            PROFILE.append((e0, e1, 2.0 * N * Ho * Wo * cout * cin * kh * kw, (N, H, W, cin, cout, kh, kw)))
    _native.check(rc)
    if pair is not None:
        if f32_out is None:
            f32_out = pair_only(shape, pair)
        else:
            f32_out._pair = pair
    if pair2 is not None:
        if f32_out2 is None:
            f32_out2 = pair_only(shape, pair2)
        else:
            f32_out2._pair = pair2
    return (f32_out, f32_out2) if alloc2 else f32_out


def linear(x2d, w_packed, cout, **kw):
    """[M,K] bf16 @ W[cout,K]^T -> [M,cout]: the dense layers of get_head (net/xception_body.py:540-558)."""
    M, K = x2d.shape
    out = conv2d_nhwc(x2d.reshape(1, 1, M, K), w_packed, cout, 1, 1, **kw)
    return out.reshape(M, cout) if out.dim() == 4 and out.shape[1] == 1 else out


def conv2d_image_fold(image_nchw_f32, w_fold, cout, kh, kw, stride, pad, **kw_args):
    """Few-channel strided convolution of the fp32 NCHW input image (the 7x7/s2 ResNet stem with explicit
    ``fixed_padding`` = ``pad`` on both sides, or Xception's 3x3/s2 VALID block1_conv1 with pad 0): one layout
    pass (image -> row-padded NHWC8 bf16) + the fold_w implicit GEMM.  Returns NHWC bf16 [N,Ho,Wo,cout]."""
    from .layout import image_to_nhwc8
    N, C, H, W = image_nchw_f32.shape
    Ho = (H + 2 * pad - kh) // stride + 1
    Wo = (W + 2 * pad - kw) // stride + 1
    if PRECISION == "fp32x3":  # generic strided convolution on the NHWC view of the fp32 image
        return conv2d_nhwc(image_nchw_f32.permute(0, 2, 3, 1), w_fold, cout, kh, kw, padding=(pad, pad, Ho, Wo),
                           strides=(stride, stride), cin=C, **kw_args)
    wp = max((Wo - 1) * stride + 8, W + pad)
    wp = (wp + 7) // 8 * 8
    if isinstance(w_fold, PairWeight):  # row-padded NHWC8 f16x2 planes of the image + the fold_w implicit GEMM
        x8 = split2(image_nchw_f32.permute(0, 2, 3, 1), cin=C, pad_w=(wp, pad))
        return conv2d_nhwc(x8, w_fold, cout, kh, kw, padding=(pad, pad, Ho, Wo), strides=(stride, stride), cin=C,
                           fold_w=(W, pad), **kw_args)
    x8 = image_to_nhwc8(image_nchw_f32, pad, wp)
    return conv2d_nhwc(x8, w_fold, cout, kh, kw, padding=(pad, pad, Ho, Wo), strides=(stride, stride), cin=C,
                       fold_w=(W, pad), **kw_args)


# ---- training: gradients of the convolution -----------------------------------------------------------------
class WgradDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("N", "H", "W", "Cin", "in_cs", "Cout", "KH", "KW", "dil_h", "dil_w",
                                            "pad_top", "pad_left", "stride_h", "stride_w", "Hout", "Wout", "dy_cs")] + \
               [("dw", ctypes.c_void_p), ("splits", ctypes.c_int), ("fold_w", ctypes.c_int), ("in_wp", ctypes.c_int)]


def conv2d_wgrad(x, dy, kh, kw, *, dilation=(1, 1), padding="SAME", strides=(1, 1), cin=None, cout=None, dw=None,
                 splits=0, fold_w=None):
    """dW of ``conv2d_nhwc(x, w, ...) -> y`` given ``dy`` (both NHWC bf16): fp32 [Cout, kh*kw, ceil(Cin/64)*64], the
    packed layout of the forward weights.  ``dw`` (zero-filled, or holding a partial sum) is accumulated into."""
    if dy.dtype == torch.float32:  # the fp32-accurate training mode
        return _conv2d_wgrad_f16x2(x, dy, kh, kw, dilation=dilation, padding=padding, strides=strides, cin=cin, cout=cout,
                                   dw=dw, splits=splits, fold_w=fold_w)
    assert x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16 and x.is_contiguous() and dy.is_contiguous()
    N, H, W, cs = x.shape
    in_wp = 0
    if fold_w is not None:  # x = row-padded NHWC8 image [N,H,in_wp,8]; fold_w = (true width, pad_left)
        in_wp, W = W, fold_w[0]
    _, Ho, Wo, dcs = dy.shape
    cin = cs if cin is None else cin
    cout = dcs if cout is None else cout
    dh, dw_ = dilation
    sh, sw = strides
    if padding == "SAME":
        pt, pl = same_pad(H, kh, dh, sh), same_pad(W, kw, dw_, sw)
    elif padding == "VALID":
        pt = pl = 0
    else:
        pt, pl = padding[:2]
    cpad = (cin + 63) // 64 * 64
    if dw is None:
        dw = torch.zeros((cout, kh, 64) if fold_w is not None else (cout, kh * kw, cpad), dtype=torch.float32,
                         device=x.device)
    d = WgradDesc(N, H, W, cin, cs, cout, kh, kw, dh, dw_, pt, pl, sh, sw, Ho, Wo, dcs, dw.data_ptr(), splits,
                  0 if fold_w is None else 1, in_wp)
    with torch.cuda.device(x.device):
        if PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = _native.lib().xdet_conv2d_wgrad_bf16(x.data_ptr(), dy.data_ptr(), ctypes.byref(d),
                                                  torch.cuda.current_stream().cuda_stream)
        if PROFILE is not None:
            e1.record()
            PROFILE.append((e0, e1, 2.0 * N * Ho * Wo * cout * cin * kh * kw, ("wgrad", N, H, W, cin, cout, kh, kw)))
    _native.check(rc)
    return dw


def _conv2d_wgrad_f16x2(x, dy, kh, kw, *, dilation, padding, strides, cin, cout, dw, splits, fold_w):
    """``conv2d_wgrad`` in "f16x2" precision: x = fp32 NHWC activation (its split planes attached by the producer, or made
    here) -- or, fold_w mode, the row-padded image planes themselves; dy = fp32 NHWC [N,Ho,Wo,>=cout]."""
    if fold_w is not None:
        xp = x  # [2,N,H,in_wp,8]
        _, N, H, in_wp, cs = xp.shape
        W = fold_w[0]
        cin = cs if cin is None else cin
    else:
        N, H, W, C = x.shape
        cin = C if cin is None else cin
        xp = pair_of(x, cin)
        cs, in_wp = xp.shape[-1], 0
    _, Ho, Wo, dC = dy.shape
    cout = dC if cout is None else cout
    dyp = split2(dy)
    dh, dw_ = dilation
    sh, sw = strides
    if padding == "SAME":
        pt, pl = same_pad(H, kh, dh, sh), same_pad(W, kw, dw_, sw)
    elif padding == "VALID":
        pt = pl = 0
    else:
        pt, pl = padding[:2]
    cpad = (cin + 63) // 64 * 64
    if dw is None:
        dw = torch.zeros((cout, kh, 64) if fold_w is not None else (cout, kh * kw, cpad), dtype=torch.float32,
                         device=dy.device)
    d = WgradDesc(N, H, W, cin, cs, cout, kh, kw, dh, dw_, pt, pl, sh, sw, Ho, Wo, dyp.shape[-1], dw.data_ptr(), splits,
                  0 if fold_w is None else 1, in_wp)
    with torch.cuda.device(dy.device):
        if PROFILE is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = _native.lib().xdet_conv2d_wgrad_f16x2(xp.data_ptr(), xp.stride(0), dyp.data_ptr(), dyp.stride(0),
                                                   ctypes.byref(d), 1.0, torch.cuda.current_stream().cuda_stream)
        if PROFILE is not None:
            e1.record()
            PROFILE.append((e0, e1, 2.0 * N * Ho * Wo * cout * cin * kh * kw, ("wgrad", N, H, W, cin, cout, kh, kw)))
    _native.check(rc)
    return dw


def pack_dgrad_weight(w_oihw):
    """Packed weights of the INPUT-gradient convolution: dX = conv(dY, flip(W) with in/out channels swapped).
    [Cout, Cin, KH, KW] -> bf16 [Cin, KH*KW*ceil(Cout/64)*64]."""
    return pack_conv_weight(torch.flip(w_oihw, dims=(2, 3)).permute(1, 0, 2, 3))


def conv2d_dgrad(dy, w_dgrad, cin, kh, kw, in_hw, *, dilation=(1, 1), padding="SAME", strides=(1, 1), cout=None, **kws):
    """dX of ``conv2d_nhwc`` (NHWC bf16): the forward kernel on ``dy`` with ``pack_dgrad_weight`` weights and the
    mirrored padding; stride-2 layers first zero-stuff ``dy`` to the input grid."""
    H, W = in_hw
    dh, dw_ = dilation
    sh, sw = strides
    if padding == "SAME":
        pt, pl = same_pad(H, kh, dh, sh), same_pad(W, kw, dw_, sw)
    elif padding == "VALID":
        pt = pl = 0
    else:
        pt, pl = padding[:2]
    if (sh, sw) != (1, 1):
        N, Ho, Wo, C = dy.shape
        up = torch.zeros((N, Ho * sh, Wo * sw, C), dtype=dy.dtype, device=dy.device)
        up[:, ::sh, ::sw] = dy  # plumbing (6 layers of the net); a fused scatter store is the obvious next step
        dy = up
    return conv2d_nhwc(dy, w_dgrad, cin, kh, kw, dilation=dilation,
                       padding=((kh - 1) * dh - pt, (kw - 1) * dw_ - pl, H, W), cin=cout, **kws)
