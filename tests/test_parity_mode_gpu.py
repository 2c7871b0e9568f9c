"""GPU: the "fp32x3" PARITY MODE (fp32 activations, split-operand tensor-core convolutions; ops/conv.py,
csrc/parity_ops.cu) against the fp32 CPU oracle, END TO END and without injected intermediates.

north_star's floating-point bar: box / score deltas within 1e-4 of the reference's fp32 path; proposal selection
bit-exact.  The bf16 throughput path cannot hold that over ~55 layers (tests/test_model_gpu.py bounds it stage by
stage); this mode can, through the same tcgen05 kernel.  Tolerances below are written per stage.
"""
import numpy as np
import pytest
import torch

from oracle import net as onet
from oracle import proposals as op

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


@pytest.fixture(scope="module")
def xd():
    assert torch.cuda.is_available()
    import xdet_b200  # noqa: F401
    from xdet_b200 import _native, ops
    _native.lib()
    return ops


def test_split3_reconstructs_fp32(xd):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn((2, 5, 7, 13), generator=g, device="cuda") * 3
    x[0, 0, 0, :4] = torch.tensor([0.0, 1e-30, -65504.0, 3.0e38], device="cuda")
    s = xd.split3(x)
    assert s.shape == (2, 5, 7, 80) and s.dtype == torch.bfloat16  # 6*13 = 78 -> 80
    C = 13
    mid, lo, hi = s[..., :C].float(), s[..., C:2 * C].float(), s[..., 2 * C:3 * C].float()
    assert torch.equal(s[..., 3 * C:4 * C].float(), mid) and torch.equal(s[..., 4 * C:5 * C].float(), hi)
    assert torch.equal(s[..., 5 * C:6 * C].float(), hi) and not s[..., 6 * C:].float().any()
    rec = hi.double() + mid.double() + lo.double()
    err = (rec - x.double()).abs()
    assert bool((err <= x.double().abs() * 2.0 ** -23 + 1e-38).all())
    # the host-side weight split is the same decomposition
    h2, m2, l2 = xd.split3_values(x)
    assert torch.equal(h2, hi) and torch.equal(m2, mid) and torch.equal(l2, lo)
    # strided (NCHW-backed) input
    xn = x.permute(0, 3, 1, 2).contiguous()
    assert torch.equal(xd.split3(xn.permute(0, 2, 3, 1)), s)


@pytest.mark.parametrize("case", [
    dict(cin=64, cout=96, k=3, stride=1, dil=1),
    dict(cin=3, cout=64, k=7, stride=2, dil=1),     # the stem as a generic strided conv over 18 channels
    dict(cin=130, cout=50, k=1, stride=1, dil=1),
    dict(cin=32, cout=40, k=3, stride=1, dil=2),
    dict(cin=48, cout=64, k=3, stride=2, dil=1),
])
def test_conv_fp32x3_vs_fp64(xd, case):
    cin, cout, k, stride, dil = case["cin"], case["cout"], case["k"], case["stride"], case["dil"]
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    x = torch.randn((2, 21, 19, cin), generator=g, device="cuda")
    w = torch.randn((cout, cin, k, k), generator=g, device="cuda") / (cin * k * k) ** 0.5
    scale = torch.rand(cout, generator=g, device="cuda") + 0.5
    bias = torch.randn(cout, generator=g, device="cuda")
    res = None
    with xd.precision("fp32x3"):
        wp = xd.pack_conv_weight(w)
        if stride == 1:
            Ho, Wo = 21, 19
            res = torch.randn((2, Ho, Wo, cout), generator=g, device="cuda")
            y = xd.conv2d_nhwc(x, wp, cout, k, k, dilation=(dil, dil), padding="SAME", scale=scale, bias=bias,
                               relu=True, residual=res)
        else:
            pad = (k - 1) // 2
            Ho, Wo = (21 + 2 * pad - k) // stride + 1, (19 + 2 * pad - k) // stride + 1
            y = xd.conv2d_nhwc(x, wp, cout, k, k, padding=(pad, pad, Ho, Wo), strides=(stride, stride), scale=scale,
                               bias=bias, relu=True)
    torch.cuda.synchronize()
    assert y.dtype == torch.float32 and y.shape == (2, Ho, Wo, cout)
    xc, wc = x.double().cpu().permute(0, 3, 1, 2), w.double().cpu()
    if stride == 1:
        pad = dil * (k - 1) // 2
        ref = torch.nn.functional.conv2d(xc, wc, padding=pad, dilation=dil)
    else:
        ref = torch.nn.functional.conv2d(xc, wc, padding=(k - 1) // 2, stride=stride)
    ref = ref * scale.double().cpu().view(1, -1, 1, 1) + bias.double().cpu().view(1, -1, 1, 1)
    if res is not None:
        ref = ref + res.double().cpu().permute(0, 3, 1, 2)
    ref = torch.relu(ref).permute(0, 2, 3, 1)
    err = (y.double().cpu() - ref).abs().max().item()
    # fp32-level: the tensor core's fp32 accumulator truncates, so the error grows ~K/16 * 2^-24 (K = 6*cin*k*k);
    # bf16 operands give ~1e-2 on the same problem
    assert err < 1e-5 * max(1.0, ref.abs().max().item()), err


def test_elementwise_fp32_forms(xd):
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((2, 10, 12, 24), generator=g, device="cuda")
    sc, bi = torch.rand(24, generator=g, device="cuda") + 0.5, torch.randn(24, generator=g, device="cuda")
    assert torch.allclose(xd.affine_relu(x, sc, bi), torch.relu(x * sc + bi), atol=1e-6)
    mp = xd.maxpool3x3s2_same(x)
    ref = torch.nn.functional.max_pool2d(torch.nn.functional.pad(x.permute(0, 3, 1, 2), (0, 1, 0, 1), value=-1e30), 3, 2)
    assert torch.equal(mp, ref.permute(0, 2, 3, 1).contiguous())  # SAME on even sizes pads the high side only
    w9 = torch.randn((9, 24), generator=g, device="cuda")
    for dil in (1, 2):
        dw = xd.depthwise3x3(x, w9, dilation=dil, relu_in=True)
        ref = torch.nn.functional.conv2d(torch.relu(x).permute(0, 3, 1, 2), w9.t().reshape(24, 1, 3, 3), padding=dil,
                                         dilation=dil, groups=24).permute(0, 2, 3, 1)
        assert torch.allclose(dw, ref, atol=1e-5)


def _run(backbone, seed):
    from xdet_b200 import light_head_rfcn_eval as lh
    params = lh.make_params(train_image_size=160, rpn_pre_nms_top_n=600, rpn_post_nms_top_n=100,
                            rpn_min_size=16.0 / 160, backbone=backbone, precision="fp32x3")
    model = lh.LightHeadRFCN(params, seed=seed)
    g = torch.Generator(device="cuda").manual_seed(seed + 100)
    images = torch.rand((2, 3, 160, 160), generator=g, device="cuda") * 2 - 1
    keys = torch.rand((2, 100), generator=g, device="cuda")
    out = model(images, shuffle_keys=keys)
    torch.cuda.synchronize()
    fm = out["rpn_feat_map"].shape[1]
    anchors = op.layer_anchors((160, 160), (fm, fm), [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5], 16)
    ref = onet.model(images.cpu().numpy(), model.store.state_dict(), params, anchors, shuffle_keys=keys.cpu().numpy())
    return out, ref


@pytest.mark.parametrize("backbone,seed", [("resnet50", 3), ("xception", 5)])
def test_end_to_end_within_1e4_of_fp32_oracle(xd, backbone, seed):
    out, ref = _run(backbone, seed)
    assert out["rpn_feat_map"].dtype == torch.float32
    # every conv stage: fp32-level agreement (relative to the tensor's max magnitude)
    assert rel(out["rpn_feat_map"].permute(0, 3, 1, 2).cpu().numpy(), ref["rpn_feat_map"]) < 1e-4
    assert rel(out["backbone_feat"].permute(0, 3, 1, 2).cpu().numpy(), ref["backbone_feat"]) < 1e-4
    assert rel(out["large_sep_feature"].cpu().numpy(), ref["large_sep_feature"]) < 1e-4
    rpn = out["rpn_out"].cpu().numpy()
    assert rel(rpn[..., :44], ref["rpn_cls"]) < 1e-4 and rel(rpn[..., 44:], ref["rpn_box"]) < 1e-4
    assert np.abs(out["rpn_object_score"].cpu().numpy() - ref["rpn_object_score"]).max() < 1e-4
    # proposal selection on the product's OWN scores equals the oracle's selection on the oracle's scores:
    # same boxes in the same order (coordinates agree to fp32 noise, so compare with a tolerance, not bits)
    assert np.abs(out["proposals_bboxes"].cpu().numpy() - ref["proposals_bboxes"]).max() < 1e-4
    # north_star: fp32 box / score deltas within 1e-4
    assert np.abs(out["cls_score"].cpu().numpy().reshape(-1, 21) - ref["cls_score"]).max() < 1e-4 * max(
        1.0, np.abs(ref["cls_score"]).max())
    assert np.abs(out["bboxes_reg"].cpu().numpy().reshape(-1, 4) - ref["bboxes_reg"]).max() < 1e-4
    assert np.abs(out["head_cls_score"].cpu().numpy() - ref["head_cls_score"]).max() < 1e-4
    assert np.abs(out["bboxes_predict"].cpu().numpy() - ref["bboxes_predict"]).max() < 1e-4
