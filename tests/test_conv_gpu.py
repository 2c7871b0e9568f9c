"""GPU: tcgen05 implicit-GEMM convolution vs a plain PyTorch fp32 reference of the same op on the
same bf16-rounded operands (fp32 accumulation on both sides).  Tolerance: the only differences are
fp32 summation order and the final bf16 rounding of the output."""
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


def ref_conv(x_nhwc, w_oihw, dil, pad_tl, out_hw):
    """fp32 conv with explicit TF-style (asymmetric) padding on bf16-rounded operands."""
    x = x_nhwc.float().permute(0, 3, 1, 2)
    w = w_oihw.to(torch.bfloat16).float()
    kh, kw = w.shape[2:]
    Ho, Wo = out_hw
    pt, pl = pad_tl
    need_h = Ho + (kh - 1) * dil[0]
    need_w = Wo + (kw - 1) * dil[1]
    pb, pr = need_h - x.shape[2] - pt, need_w - x.shape[3] - pl
    x = F.pad(x, (pl, max(pr, 0), pt, max(pb, 0)))
    return F.conv2d(x, w, dilation=dil)[:, :, :Ho, :Wo]


CASES = [
    # N, H, W, Cin, Cout, KH, KW, dil
    (2, 30, 30, 64, 128, 1, 1, 1),
    (1, 30, 30, 256, 64, 3, 3, 1),
    (2, 30, 30, 128, 256, 3, 3, 2),     # dilated (layer4 / xception exit flow)
    (1, 30, 30, 192, 256, 15, 1, 1),    # large separable, vertical
    (1, 30, 30, 256, 496, 1, 15, 1),    # large separable, horizontal, Cout not a multiple of 128
    (1, 60, 60, 128, 128, 3, 3, 1),
    (1, 120, 120, 64, 64, 3, 3, 1),
    (1, 50, 50, 728, 728, 1, 1, 1),     # Cin not a multiple of 64
    (3, 17, 23, 72, 40, 3, 3, 1),       # ragged everything
]


@pytest.mark.parametrize("case", CASES)
def test_conv_matches_fp32_reference(ops, case):
    N, H, W, Cin, Cout, KH, KW, dil = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    x = torch.randn((N, H, W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn((Cout, Cin, KH, KW), generator=g, device="cuda") / (Cin * KH * KW) ** 0.5
    wp = ops.pack_conv_weight(w)
    y = ops.conv2d_nhwc(x, wp, Cout, KH, KW, dilation=(dil, dil), out_layout="nhwc_f32")
    torch.cuda.synchronize()
    ref = ref_conv(x, w, (dil, dil), (ops.same_pad(H, KH, dil), ops.same_pad(W, KW, dil)), (H, W)).permute(0, 2, 3, 1)
    err = (y - ref).abs().max().item()
    assert err < 2e-3, err


def test_epilogue_scale_bias_relu_residual_dual_output(ops):
    N, H, W, Cin, Cout = 2, 30, 30, 256, 256
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.randn((N, H, W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn((Cout, Cin, 1, 1), generator=g, device="cuda") / Cin ** 0.5
    scale = torch.rand(Cout, generator=g, device="cuda") + 0.5
    bias = torch.randn(Cout, generator=g, device="cuda")
    scale2 = torch.rand(Cout, generator=g, device="cuda") + 0.5
    bias2 = torch.randn(Cout, generator=g, device="cuda")
    res = torch.randn((N, H, W, Cout), generator=g, device="cuda").to(torch.bfloat16)
    out2 = torch.empty((N, H, W, Cout), dtype=torch.bfloat16, device="cuda")
    y = ops.conv2d_nhwc(x, ops.pack_conv_weight(w), Cout, 1, 1, scale=scale, bias=bias, residual=res, out2=out2,
                        scale2=scale2, bias2=bias2)
    torch.cuda.synchronize()
    acc = ref_conv(x, w, (1, 1), (0, 0), (H, W)).permute(0, 2, 3, 1)
    v = acc * scale + bias + res.float()
    assert (y.float() - v).abs().max().item() < 0.05  # bf16 output rounding (|v| up to ~6)
    v2 = torch.relu(v * scale2 + bias2)
    assert (out2.float() - v2).abs().max().item() < 0.06
    y3 = ops.conv2d_nhwc(x, ops.pack_conv_weight(w), Cout, 1, 1, bias=bias, relu=True, out_layout="nchw_f32")
    torch.cuda.synchronize()
    assert (y3 - torch.relu(acc + bias).permute(0, 3, 1, 2)).abs().max().item() < 2e-3


def test_valid_padding_and_linear(ops):
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((1, 20, 20, 64), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn((64, 64, 3, 3), generator=g, device="cuda") / 24.0
    y = ops.conv2d_nhwc(x, ops.pack_conv_weight(w), 64, 3, 3, padding="VALID", out_layout="nhwc_f32")
    ref = ref_conv(x, w, (1, 1), (0, 0), (18, 18)).permute(0, 2, 3, 1)
    assert y.shape == (1, 18, 18, 64) and (y - ref).abs().max().item() < 2e-3
    a = torch.randn((1000, 496), generator=g, device="cuda").to(torch.bfloat16)
    wl = torch.randn((2048, 496), generator=g, device="cuda") / 22.0
    yl = ops.linear(a, ops.pack_conv_weight(wl.reshape(2048, 496, 1, 1)), 2048, out_layout="nhwc_f32")
    refl = a.float() @ wl.to(torch.bfloat16).float().t()
    assert (yl.reshape(1000, 2048) - refl).abs().max().item() < 2e-3



def ref_conv_strided(x_nhwc, w_oihw, stride, pad):
    x = x_nhwc.float().permute(0, 3, 1, 2)
    w = w_oihw.to(torch.bfloat16).float()
    return F.conv2d(x, w, stride=stride, padding=pad)


STRIDED = [
    # N, H, W, Cin, Cout, K, stride, pad   (conv2d_fixed_padding with strides 2: explicit pad (k-1)//2, then VALID)
    (2, 60, 60, 128, 128, 3, 2, 1),
    (2, 120, 120, 256, 512, 1, 2, 0),
    (1, 61, 47, 64, 96, 3, 2, 1),     # odd sizes
    (1, 59, 59, 32, 64, 3, 2, 0),     # VALID (xception block1_conv1-like, on NHWC input)
]


@pytest.mark.parametrize("case", STRIDED)
def test_strided_conv_matches_fp32_reference(ops, case):
    N, H, W, Cin, Cout, K, s, pad = case
    g = torch.Generator(device="cuda").manual_seed(sum(case))
    x = torch.randn((N, H, W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn((Cout, Cin, K, K), generator=g, device="cuda") / (Cin * K * K) ** 0.5
    ref = ref_conv_strided(x, w, s, pad).permute(0, 2, 3, 1)
    Ho, Wo = ref.shape[1:3]
    y = ops.conv2d_nhwc(x, ops.pack_conv_weight(w), Cout, K, K, padding=(pad, pad, Ho, Wo), strides=(s, s),
                        out_layout="nhwc_f32")
    torch.cuda.synchronize()
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() < 2e-3
    yb = ops.conv2d_nhwc(x, ops.pack_conv_weight(w), Cout, K, K, padding=(pad, pad, Ho, Wo), strides=(s, s))
    torch.cuda.synchronize()
    assert (yb.float() - ref).abs().max().item() < 0.03  # bf16 output rounding


@pytest.mark.parametrize("shape", [(2, 3, 96, 96, 64, 7, 2, 3), (1, 3, 480, 480, 64, 7, 2, 3), (2, 3, 95, 131, 32, 3, 2, 0)])
def test_image_fold_conv(ops, shape):
    """The few-channel stem convolution (fold_w mode) on the fp32 NCHW image."""
    N, C, H, W, Cout, K, s, pad = shape
    g = torch.Generator(device="cuda").manual_seed(sum(shape))
    img = torch.rand((N, C, H, W), generator=g, device="cuda") * 2 - 1
    w = torch.randn((Cout, C, K, K), generator=g, device="cuda") / (C * K * K) ** 0.5
    y = ops.conv2d_image_fold(img, ops.pack_fold_weight(w), Cout, K, K, s, pad, out_layout="nhwc_f32")
    torch.cuda.synchronize()
    ref = F.conv2d(img.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), stride=s, padding=pad).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    assert (y - ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("epi_groups", [1, 2])
@pytest.mark.parametrize("block_n", [0, 64, 128, 256])
def test_persistent_many_tiles_residual_chain(ops, block_n, epi_groups):
    """More tiles than SMs (each CTA walks several tiles): accumulator double buffering, the residual prefetch
    chain across tiles, clipped channel tail (728 = 11*64 + 24), every N-tile width."""
    N, H, W, Cin, Cout = 8, 60, 60, 128, 728
    g = torch.Generator(device="cuda").manual_seed(11 + block_n)
    x = torch.randn((N, H, W, Cin), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn((Cout, Cin, 1, 1), generator=g, device="cuda") / Cin ** 0.5
    scale = torch.rand(Cout, generator=g, device="cuda") + 0.5
    bias = torch.randn(Cout, generator=g, device="cuda")
    scale2 = torch.rand(Cout, generator=g, device="cuda") + 0.5
    bias2 = torch.randn(Cout, generator=g, device="cuda")
    res = torch.randn((N, H, W, Cout), generator=g, device="cuda").to(torch.bfloat16)
    out2 = torch.full((N, H, W, Cout), 7.0, dtype=torch.bfloat16, device="cuda")
    y = ops.conv2d_nhwc(x, ops.pack_conv_weight(w), Cout, 1, 1, scale=scale, bias=bias, residual=res, out2=out2,
                        scale2=scale2, bias2=bias2, block_n=block_n, epi_groups=epi_groups)
    torch.cuda.synchronize()
    acc = ref_conv(x, w, (1, 1), (0, 0), (H, W)).permute(0, 2, 3, 1)
    v = acc * scale + bias + res.float()
    assert (y.float() - v).abs().max().item() < 0.05
    assert (out2.float() - torch.relu(v * scale2 + bias2)).abs().max().item() < 0.06
    # second output only (raw sum not stored), and plain (no residual) with two groups
    o2 = torch.full((N, H, W, Cout), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.conv2d_nhwc(x, ops.pack_conv_weight(w), Cout, 1, 1, scale=scale, bias=bias, residual=res, out2=o2,
                    scale2=scale2, bias2=bias2, block_n=block_n, epi_groups=epi_groups, skip_out=True)
    yp = ops.conv2d_nhwc(x, ops.pack_conv_weight(w), Cout, 1, 1, scale=scale, bias=bias, relu=True, block_n=block_n,
                         epi_groups=epi_groups)
    torch.cuda.synchronize()
    assert torch.equal(o2, out2)
    assert (yp.float() - torch.relu(acc * scale + bias)).abs().max().item() < 0.05
    # spatial (non-flattened) tiles with ragged edges + residual: 3x3 on 30x30
    w3 = torch.randn((256, Cin, 3, 3), generator=g, device="cuda") / (9 * Cin) ** 0.5
    x3 = x[:, :30, :30].contiguous()
    r3 = torch.randn((N, 30, 30, 256), generator=g, device="cuda").to(torch.bfloat16)
    y3 = ops.conv2d_nhwc(x3, ops.pack_conv_weight(w3), 256, 3, 3, residual=r3, relu=True, block_n=block_n,
                         epi_groups=epi_groups)
    torch.cuda.synchronize()
    ref3 = torch.relu(ref_conv(x3, w3, (1, 1), (1, 1), (30, 30)).permute(0, 2, 3, 1) + r3.float())
    assert (y3.float() - ref3).abs().max().item() < 0.05


def test_deep_k_and_wide_n(ops):
    """large_sep-like: K = 15*2048 with BN = 256 tiles; RPN-like fp32 NHWC output with 132 channels."""
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((2, 30, 30, 2048), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn((512, 2048, 15, 1), generator=g, device="cuda") / (15 * 2048) ** 0.5
    for bn in (0, 128, 256):
        y = ops.conv2d_nhwc(x, ops.pack_conv_weight(w), 512, 15, 1, out_layout="nhwc_f32", block_n=bn)
        torch.cuda.synchronize()
        ref = ref_conv(x, w, (1, 1), (7, 0), (30, 30)).permute(0, 2, 3, 1)
        assert (y - ref).abs().max().item() < 3e-3, bn
    w2 = torch.randn((132, 2048, 1, 1), generator=g, device="cuda") / 2048 ** 0.5
    y2 = ops.conv2d_nhwc(x, ops.pack_conv_weight(w2), 132, 1, 1, out_layout="nhwc_f32")
    ref2 = ref_conv(x, w2, (1, 1), (0, 0), (30, 30)).permute(0, 2, 3, 1)
    assert (y2 - ref2).abs().max().item() < 2e-3


@pytest.mark.parametrize("case", [(2, 30, 30, 728, 1, True), (1, 77, 77, 64, 1, False), (2, 25, 31, 1024, 2, False),
                                  (1, 39, 39, 256, 1, True)])
def test_depthwise3x3(ops, case):
    """Depthwise half of tf.layers.separable_conv2d (SAME, depth multiplier 1) vs F.conv2d(groups=C) in fp32."""
    N, H, W, C, dil, relu_in = case
    g = torch.Generator(device="cuda").manual_seed(sum(map(int, case)))
    x = torch.randn((N, H, W, C), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn((3, 3, C), generator=g, device="cuda") / 3.0
    y = ops.depthwise3x3(x, w.reshape(9, C).contiguous(), dilation=dil, relu_in=relu_in)
    torch.cuda.synchronize()
    xi = x.float().permute(0, 3, 1, 2)
    if relu_in:
        xi = torch.relu(xi)
    ref = F.conv2d(xi, w.permute(2, 0, 1).unsqueeze(1), padding=dil, dilation=dil, groups=C).permute(0, 2, 3, 1)
    assert (y.float() - ref).abs().max().item() < 0.03  # bf16 output rounding


def test_maxpool_add(ops):
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn((2, 77, 77, 128), generator=g, device="cuda").to(torch.bfloat16)
    r = torch.randn((2, 39, 39, 128), generator=g, device="cuda").to(torch.bfloat16)
    y = ops.maxpool3x3s2_same(x, residual=r)
    torch.cuda.synchronize()
    xp = F.pad(x.float().permute(0, 3, 1, 2), (1, 1, 1, 1), value=float("-inf"))  # 77 -> 39: pad (1,1)
    ref = F.max_pool2d(xp, 3, 2).permute(0, 2, 3, 1) + r.float()
    assert (y.float() - ref).abs().max().item() < 0.03
