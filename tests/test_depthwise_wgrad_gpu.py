"""GPU: numerics of the depthwise 3x3 weight-gradient kernel (csrc/depthwise_wgrad.cu) and of the 'input gradient =
forward kernel with flipped taps' identity, against torch autograd in fp64 on the same bf16-rounded operands."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    assert torch.cuda.is_available()
    import xdet_b200  # noqa: F401
    from xdet_b200 import _native
    return _native.lib()


@pytest.mark.parametrize("shape,dil,relu_in", [((2, 19, 23, 72), 1, True), ((1, 50, 50, 728), 1, True),
                                               ((3, 10, 12, 1536), 2, False), ((1, 5, 4, 8), 2, True),
                                               # the row-sliding kernel's corners: widths around the ring length (3 / 5),
                                               # rows split over several warps per CTA, the training shapes
                                               ((2, 7, 3, 16), 1, False), ((2, 7, 4, 136), 1, True),
                                               ((1, 9, 5, 24), 2, True), ((1, 9, 6, 264), 2, False),
                                               ((1, 3, 2, 8), 1, True),       # W <= 2*dil: the pixel-strided fallback
                                               ((8, 30, 30, 728), 1, True), ((2, 120, 120, 128), 1, False),
                                               ((8, 30, 30, 1024), 2, True)])
def test_depthwise_wgrad_and_flipped_dgrad(lib, shape, dil, relu_in):
    from xdet_b200 import _native, ops
    N, H, W, C = shape
    g = torch.Generator(device="cuda").manual_seed(C + H)
    x = torch.randn(shape, generator=g, device="cuda").to(torch.bfloat16)
    dy = torch.randn(shape, generator=g, device="cuda").to(torch.bfloat16)
    w9 = torch.randn((9, C), generator=g, device="cuda")
    dw = torch.zeros((9, C), device="cuda")
    _native.check(lib.xdet_depthwise3x3_wgrad_bf16(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), N, H, W, C, dil,
                                                   1 if relu_in else 0, torch.cuda.current_stream().cuda_stream))
    dx = ops.depthwise3x3(dy, w9.flip(0).contiguous(), dilation=dil, relu_in=False)   # taps flipped: (kh,kw)->(2-kh,2-kw)
    torch.cuda.synchronize()
    xr = x.double().cpu().permute(0, 3, 1, 2).requires_grad_(True)
    wr = w9.double().cpu().t().reshape(C, 1, 3, 3).clone().requires_grad_(True)
    a = torch.relu(xr) if relu_in else xr
    a.retain_grad()
    y = torch.nn.functional.conv2d(a, wr, padding=dil, dilation=dil, groups=C)
    y.backward(dy.double().cpu().permute(0, 3, 1, 2))
    want_dw = wr.grad.reshape(C, 9).t()
    err = (dw.double().cpu() - want_dw).abs().max().item()
    assert err < 1e-4 * max(1.0, want_dw.abs().max().item()), err       # fp32 accumulation of exact bf16 products
    want_dx = a.grad.permute(0, 2, 3, 1)                                # gradient w.r.t. the (ReLU'd) conv input
    errx = (dx.double().cpu() - want_dx).abs().max().item()
    assert errx < 2e-2 * max(1.0, want_dx.abs().max().item()), errx      # bf16 output rounding
    assert np.isfinite(err) and np.isfinite(errx)
