"""GPU: gradients of the tensor-core convolution (wgrad: MN-major tcgen05 GEMM over the pixel axis; dgrad: the
forward kernel on dY with flipped weights) against torch autograd in fp32 on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    assert torch.cuda.is_available()
    import xdet_b200  # noqa: F401
    from xdet_b200 import ops
    return ops


CASES = [
    # N, H, W, Cin, Cout, K, stride, dil
    (2, 30, 30, 64, 128, 1, 1, 1),
    (2, 30, 30, 128, 64, 3, 1, 1),
    (1, 30, 30, 256, 256, 3, 1, 2),
    (2, 17, 23, 72, 40, 3, 1, 1),      # ragged everything
    (2, 60, 60, 128, 128, 3, 2, 1),    # stride 2, fixed padding
    (2, 60, 60, 256, 512, 1, 2, 1),
    (8, 30, 30, 512, 320, 1, 1, 1),    # several pixel splits / tiles
]


def reference(x, w, dy, K, s, dil, pad):
    xf = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wf = w.to(torch.bfloat16).float().requires_grad_(True)
    y = F.conv2d(xf, wf, stride=s, padding=pad, dilation=dil)
    y.backward(dy.float().permute(0, 3, 1, 2))
    return y, xf.grad.permute(0, 2, 3, 1), wf.grad


@pytest.mark.parametrize("case", CASES)
def test_wgrad_and_dgrad(ops, case):
    N, H, W, Cin, Cout, K, s, dil = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    x = torch.randn((N, H, W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn((Cout, Cin, K, K), generator=g, device="cuda") / (Cin * K * K) ** 0.5
    pad = dil * (K - 1) // 2
    Ho, Wo = (H + 2 * pad - dil * (K - 1) - 1) // s + 1, (W + 2 * pad - dil * (K - 1) - 1) // s + 1
    dy = torch.randn((N, Ho, Wo, Cout), generator=g, device="cuda").to(torch.bfloat16)
    y_ref, dx_ref, dw_ref = reference(x, w, dy, K, s, dil, pad)
    assert y_ref.shape[2:] == (Ho, Wo)
    padding = (pad, pad, Ho, Wo)
    dw = ops.conv2d_wgrad(x, dy, K, K, dilation=(dil, dil), padding=padding, strides=(s, s))
    torch.cuda.synchronize()
    got = dw.reshape(Cout, K, K, -1)[..., :Cin].permute(0, 3, 1, 2)
    scale = dw_ref.abs().max().item()
    assert (got - dw_ref).abs().max().item() < 2e-3 * max(scale, 1.0), (got - dw_ref).abs().max().item()
    assert dw.reshape(Cout, K * K, -1)[..., Cin:].abs().max().item() == 0 if dw.shape[-1] > Cin else True
    dx = ops.conv2d_dgrad(dy, ops.pack_dgrad_weight(w), Cin, K, K, (H, W), dilation=(dil, dil), padding=padding,
                          strides=(s, s), out_layout="nhwc_f32")
    torch.cuda.synchronize()
    assert dx.shape == dx_ref.shape
    assert (dx - dx_ref).abs().max().item() < 2e-3 * max(dx_ref.abs().max().item(), 1.0)


def test_wgrad_accumulates_and_splits_agree(ops):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn((4, 30, 30, 128), generator=g, device="cuda").to(torch.bfloat16)
    dy = torch.randn((4, 30, 30, 128), generator=g, device="cuda").to(torch.bfloat16)
    a = ops.conv2d_wgrad(x, dy, 3, 3, splits=1)
    b = ops.conv2d_wgrad(x, dy, 3, 3, splits=7)
    c = ops.conv2d_wgrad(x, dy, 3, 3, dw=a.clone())
    torch.cuda.synchronize()
    assert (a - b).abs().max().item() < 1e-3 * a.abs().max().item()
    assert (c - 2 * a).abs().max().item() < 1e-3 * a.abs().max().item()


def test_stem_fold_wgrad(ops):
    """dW of the 7x7/s2 stem on the 3-channel image (fold_w operand: a K chunk = one filter row of the padded NHWC8 image)."""
    g = torch.Generator(device="cuda").manual_seed(8)
    N, H, W = 2, 96, 96
    img = torch.rand((N, 3, H, W), generator=g, device="cuda") * 2 - 1
    w = torch.randn((64, 3, 7, 7), generator=g, device="cuda") / 147 ** 0.5
    Ho = Wo = 48
    wp = (max((Wo - 1) * 2 + 8, W + 3) + 7) // 8 * 8
    x8 = ops.image_to_nhwc8(img, 3, wp)
    dy = torch.randn((N, Ho, Wo, 64), generator=g, device="cuda").to(torch.bfloat16)
    dw = ops.conv2d_wgrad(x8, dy, 7, 7, padding=(3, 3, Ho, Wo), strides=(2, 2), cin=3, cout=64, fold_w=(W, 3))
    torch.cuda.synchronize()
    xf = img.to(torch.bfloat16).float()
    wf = w.clone().requires_grad_(True)
    F.conv2d(xf, wf, stride=2, padding=3).backward(dy.float().permute(0, 3, 1, 2))
    got = dw.reshape(64, 7, 8, 8)[:, :, :7, :3].permute(0, 3, 1, 2)  # [co][kh][kw][c] -> [co][c][kh][kw]
    assert (got - wf.grad).abs().max().item() < 2e-3 * max(1.0, wf.grad.abs().max().item())
