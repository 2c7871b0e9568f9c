"""Host wrappers of the bandwidth helpers in ``csrc/layout_ops.cu`` (bf16 NHWC tensors)."""
import torch

from .. import _native
from .conv import same_pad


def _st():
    return torch.cuda.current_stream().cuda_stream


def im2col(x, kh, kw, stride, pad_top, pad_left, Ho, Wo, *, nchw_f32=False, cin=None, out_cs=None):
    """x: NHWC bf16 [N,H,W,cs] (or NCHW fp32 [N,C,H,W] with nchw_f32=True) -> [N,Ho,Wo,out_cs] bf16 patches,
    channel order (kh, kw, c).  See xdet_im2col_bf16."""
    if nchw_f32:
        N, C, H, W = x.shape
        cs = C
        assert x.dtype == torch.float32 and x.is_contiguous()
    else:
        N, H, W, cs = x.shape
        C = cs if cin is None else cin
        assert x.dtype == torch.bfloat16 and x.is_contiguous()
    K = kh * kw * C
    out_cs = (K + 7) // 8 * 8 if out_cs is None else out_cs
    out = torch.empty((N, Ho, Wo, out_cs), dtype=torch.bfloat16, device=x.device)
    rc = _native.lib().xdet_im2col_bf16(x.data_ptr(), 1 if nchw_f32 else 0, out.data_ptr(), N, H, W, C, cs, kh, kw,
                                        stride, pad_top, pad_left, Ho, Wo, out_cs, _st())
    _native.check(rc)
    return out


def maxpool3x3s2_same(x, scale2=None, bias2=None, residual=None, forms="both", forms2="both"):
    """tf.layers.max_pooling2d(3, 2, 'SAME') on NHWC bf16 (+ ``residual`` added to the pooled value); optionally
    also returns relu(y*scale2+bias2).  ``forms`` / ``forms2`` ("f16x2" precision only, see ops.conv2d_nhwc): which
    forms of the two results are stored -- "both", "f32", "pair", or "none" (the pooled value itself is not needed)."""
    N, H, W, C = x.shape
    Ho, Wo = -(-H // 2), -(-W // 2)
    pt, pl = same_pad(H, 3, 1, 2), same_pad(W, 3, 1, 2)
    from . import conv as _conv
    if x.dtype == torch.float32 and _conv.PRECISION == "f16x2" and C % 8 == 0:
        assert x.is_contiguous() and not getattr(x, "_pair_only", False)
        dev = x.device
        shape = (N, Ho, Wo, C)
        want2 = scale2 is not None
        out = torch.empty(shape, dtype=torch.float32, device=dev) if forms in ("both", "f32") else None
        pair = torch.empty((2,) + shape, dtype=torch.float16, device=dev) if forms in ("both", "pair") else None
        out2 = torch.empty(shape, dtype=torch.float32, device=dev) if want2 and forms2 in ("both", "f32") else None
        pair2 = torch.empty((2,) + shape, dtype=torch.float16, device=dev) if want2 and forms2 in ("both", "pair") else None
        if residual is not None:
            assert residual.shape == shape and residual.dtype == torch.float32 and residual.is_contiguous()
        some = pair if pair is not None else pair2
        p = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        rc = _native.lib().xdet_maxpool3x3s2_f32x(x.data_ptr(), p(out), p(pair), p(out2), p(pair2), p(scale2), p(bias2),
                                                  p(residual), 0 if some is None else some.stride(0), N, H, W, C, Ho, Wo,
                                                  pt, pl, _st())
        _native.check(rc)

        def wrap(f32, pr):
            if f32 is None:
                return _conv.pair_only(shape, pr) if pr is not None else torch.empty((1,), device=dev).expand(shape)
            if pr is not None:
                f32._pair = pr
            return f32
        return (wrap(out, pair), wrap(out2, pair2)) if want2 else wrap(out, pair)
    if x.dtype == torch.float32:  # parity mode (csrc/parity_ops.cu)
        assert x.is_contiguous()
        out = torch.empty((N, Ho, Wo, C), dtype=torch.float32, device=x.device)
        out2 = torch.empty_like(out) if scale2 is not None else None
        if residual is not None:
            assert residual.shape == out.shape and residual.dtype == torch.float32 and residual.is_contiguous()
        rc = _native.lib().xdet_maxpool3x3s2_f32(
            x.data_ptr(), out.data_ptr(), None if out2 is None else out2.data_ptr(),
            None if scale2 is None else scale2.data_ptr(), None if bias2 is None else bias2.data_ptr(),
            None if residual is None else residual.data_ptr(), N, H, W, C, Ho, Wo, pt, pl, _st())
        _native.check(rc)
        return (out, out2) if out2 is not None else out
    out = torch.empty((N, Ho, Wo, C), dtype=torch.bfloat16, device=x.device)
    out2 = torch.empty_like(out) if scale2 is not None else None
    if residual is not None:
        assert residual.shape == out.shape and residual.dtype == torch.bfloat16 and residual.is_contiguous()
    rc = _native.lib().xdet_maxpool3x3s2_add_bf16(
        x.data_ptr(), out.data_ptr(), None if out2 is None else out2.data_ptr(),
        None if scale2 is None else scale2.data_ptr(), None if bias2 is None else bias2.data_ptr(),
        None if residual is None else residual.data_ptr(), N, H, W, C, Ho, Wo, pt, pl, _st())
    _native.check(rc)
    return (out, out2) if out2 is not None else out


def depthwise3x3(x, w9c, dilation=1, relu_in=False, forms="pair"):
    """Depthwise 3x3 SAME conv (depth multiplier 1) on NHWC bf16; ``w9c`` = [9, C] fp32 taps (kh-major).
    ``forms`` ("f16x2" precision only): the result's forms -- its reader in XceptionBody is the pointwise convolution,
    so by default only the split planes are written."""
    N, H, W, C = x.shape
    from . import conv as _conv
    if x.dtype == torch.float32 and _conv.PRECISION == "f16x2" and C % 8 == 0:
        assert x.is_contiguous() and not getattr(x, "_pair_only", False)
        assert w9c.dtype == torch.float32 and w9c.shape == (9, C)
        out = torch.empty_like(x) if forms in ("both", "f32") else None
        pair = torch.empty((2, N, H, W, C), dtype=torch.float16, device=x.device) if forms in ("both", "pair") else None
        rc = _native.lib().xdet_depthwise3x3_f32x(x.data_ptr(), w9c.data_ptr(), None if out is None else out.data_ptr(),
                                                  None if pair is None else pair.data_ptr(),
                                                  0 if pair is None else pair.stride(0), N, H, W, C, dilation,
                                                  1 if relu_in else 0, _st())
        _native.check(rc)
        if out is None:
            return _conv.pair_only((N, H, W, C), pair)
        if pair is not None:
            out._pair = pair
        return out
    if x.dtype == torch.float32:  # parity mode (csrc/parity_ops.cu)
        assert x.is_contiguous() and w9c.dtype == torch.float32 and w9c.shape == (9, C)
        out = torch.empty_like(x)
        rc = _native.lib().xdet_depthwise3x3_f32(x.data_ptr(), w9c.data_ptr(), out.data_ptr(), N, H, W, C, dilation,
                                                 1 if relu_in else 0, _st())
        _native.check(rc)
        return out
    assert x.dtype == torch.bfloat16 and x.is_contiguous() and w9c.dtype == torch.float32 and w9c.shape == (9, C)
    out = torch.empty_like(x)
    rc = _native.lib().xdet_depthwise3x3_bf16(x.data_ptr(), w9c.data_ptr(), out.data_ptr(), N, H, W, C, dilation,
                                              1 if relu_in else 0, _st())
    _native.check(rc)
    return out


def affine_relu(x, scale, bias, relu=True):
    """Inference batch-norm (+ReLU) on NHWC bf16: y = x*scale[c] + bias[c]."""
    C = x.shape[-1]
    if x.dtype == torch.float32:  # parity mode (csrc/parity_ops.cu)
        from .conv import f32_post
        return f32_post(x, scale2=scale, bias2=bias, relu2=relu, store_out=False)[1]
    out = torch.empty_like(x)
    rc = _native.lib().xdet_affine_relu_bf16(x.data_ptr(), out.data_ptr(), scale.data_ptr(), bias.data_ptr(),
                                             x.numel() // C, C, 1 if relu else 0, _st())
    _native.check(rc)
    return out


def f32_to_bf16_rows(x2d, pitch):
    rows, cols = x2d.shape
    out = torch.empty((rows, pitch), dtype=torch.bfloat16, device=x2d.device)
    rc = _native.lib().xdet_f32_to_bf16_rows(x2d.data_ptr(), out.data_ptr(), rows, cols, pitch, _st())
    _native.check(rc)
    return out


def image_to_nhwc8(image_nchw_f32, pad_left, wp):
    """[N,C<=8,H,W] fp32 NCHW -> [N,H,wp,8] bf16, `pad_left` zero pixels in front of each row (fold_w conv input)."""
    N, C, H, W = image_nchw_f32.shape
    assert image_nchw_f32.dtype == torch.float32 and image_nchw_f32.is_contiguous()
    out = torch.empty((N, H, wp, 8), dtype=torch.bfloat16, device=image_nchw_f32.device)
    rc = _native.lib().xdet_image_to_nhwc8_bf16(image_nchw_f32.data_ptr(), out.data_ptr(), N, C, H, W, wp, pad_left,
                                                _st())
    _native.check(rc)
    return out
