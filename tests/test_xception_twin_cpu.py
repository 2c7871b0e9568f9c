"""CPU: the COMPOSITION logic of the staged device-side training backbone
(x-detector_b200/net/xception_train.py) -- argument orders, geometry tuples, NHWC / packed-gradient layouts,
the TF-layout conversions of the returned gradients, the order of the backward walk -- checked before its first GPU
run by executing it with every kernel wrapper replaced by a float64 torch-CPU stand-in that honours the wrapper's
documented contract (same signatures, NHWC bf16 tensors in and out, the packed [Cout, kh*kw, cin_pad] weight-gradient
layout, ``conv2d_dgrad`` left REAL so that its padding / zero-stuffing arithmetic is exercised through the stand-in
forward).  Reference: the CPU blueprint with bf16 storage emulated at the same points.  What this cannot see: the
kernels themselves (their own GPU tests) and any misreading of a wrapper's contract shared by stand-in and twin."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import xdet_b200  # noqa: F401
from oracle import net as onet
from oracle import xception_backward as xb
from xdet_b200 import ops
from xdet_b200.net import xception_train as xt
from xdet_b200.ops import conv as conv_mod
from xdet_b200.ops import train as T

GOLD = os.path.join(os.path.dirname(__file__), "golden", "netgraph_golden.npz")
BF = torch.bfloat16


def nchw(x, c=None):
    return (x if c is None else x[..., :c]).double().permute(0, 3, 1, 2)


def nhwc_bf16(y):
    return y.permute(0, 2, 3, 1).contiguous().to(BF)


def geometry(H, W, kh, kw, dilation, padding, strides):
    (dh, dw), (sh, sw) = dilation, strides
    if padding == "SAME":
        pt, pl = conv_mod.same_pad(H, kh, dh, sh), conv_mod.same_pad(W, kw, dw, sw)
        Ho, Wo = -(-H // sh), -(-W // sw)
    elif padding == "VALID":
        pt = pl = 0
        Ho, Wo = (H - (kh - 1) * dh - 1) // sh + 1, (W - (kw - 1) * dw - 1) // sw + 1
    else:
        pt, pl, Ho, Wo = padding
    pb = (Ho - 1) * sh + (kh - 1) * dh + 1 - H - pt
    pr = (Wo - 1) * sw + (kw - 1) * dw + 1 - W - pl
    return (pl, pr, pt, pb), Ho, Wo


def fake_pack_conv_weight(w_oihw):
    return w_oihw.double().clone()


def fake_conv2d_nhwc(x, w, cout, kh, kw, *, dilation=(1, 1), padding="SAME", strides=(1, 1), cin=None, **unused):
    assert x.dtype == BF and x.is_contiguous() and not unused, unused
    N, H, W, cs = x.shape
    cin = cs if cin is None else cin
    assert tuple(w.shape) == (cout, cin, kh, kw), (tuple(w.shape), (cout, cin, kh, kw))
    pads, Ho, Wo = geometry(H, W, kh, kw, dilation, padding, strides)
    y = F.conv2d(F.pad(nchw(x, cin), pads), w, stride=strides, dilation=dilation)
    assert y.shape[2:] == (Ho, Wo), (y.shape, Ho, Wo)
    return nhwc_bf16(y)


def fake_conv2d_wgrad(x, dy, kh, kw, *, dilation=(1, 1), padding="SAME", strides=(1, 1), cin=None, cout=None, dw=None):
    assert x.dtype == BF and dy.dtype == BF and x.is_contiguous() and dy.is_contiguous()
    N, H, W, cs = x.shape
    pads, Ho, Wo = geometry(H, W, kh, kw, dilation, padding, strides)
    assert dy.shape[1:3] == (Ho, Wo)
    xp = F.pad(nchw(x, cin), pads)
    g = torch.nn.grad.conv2d_weight(xp, (cout, cin, kh, kw), nchw(dy, cout), stride=strides, dilation=dilation)
    assert dw.shape == (cout, kh * kw, (cin + 63) // 64 * 64) and dw.dtype == torch.float32
    dw[:, :, :cin] += g.permute(0, 2, 3, 1).reshape(cout, kh * kw, cin).float()


def fake_depthwise3x3(x, w9, dilation=1, relu_in=False, forms=None):
    assert x.dtype == BF and x.is_contiguous() and w9.shape == (9, x.shape[-1]) and w9.dtype == torch.float32
    a = nchw(x)
    a = torch.relu(a) if relu_in else a
    return nhwc_bf16(xb.depthwise_fwd(a, w9.double().reshape(3, 3, -1), dilation))


def fake_depthwise3x3_wgrad(x, dy, dilation, relu_in, out=None):
    if out is not None:   # the device kernel ACCUMULATES into a view of the flat gradient buffer
        out += fake_depthwise3x3_wgrad(x, dy, dilation, relu_in).reshape(-1)
        return out
    a, g, d = nchw(x), nchw(dy), dilation
    a = torch.relu(a) if relu_in else a
    H, W = a.shape[2:]
    ap = F.pad(a, (d, d, d, d))
    dw = torch.stack([(ap[:, :, kh * d:kh * d + H, kw * d:kw * d + W] * g).sum(dim=(0, 2, 3))
                      for kh in range(3) for kw in range(3)])
    return dw.float()


def fake_affine_relu(x, scale, bias, relu=True):
    y = x.double() * scale.double() + bias.double()
    return (torch.relu(y) if relu else y).to(BF)


def fake_bn_train(x, gamma, beta, eps, decay=None, moving_mean=None, moving_var=None):
    x2 = x.double().reshape(-1, x.shape[-1])
    st = T.BNState()
    st.rows = x2.shape[0]
    st.mean = x2.mean(0)
    st.invstd = 1.0 / torch.sqrt(x2.var(0, unbiased=False) + eps)
    st.scale = gamma.double() * st.invstd
    st.shift = beta.double() - st.mean * st.scale
    return st


def fake_bn_relu_bwd(dy, x, st, relu=True, add_in=None):
    assert dy.shape == x.shape and dy.dtype == BF and dy.is_contiguous() and add_in is None
    C = x.shape[-1]
    g, xd = dy.double().reshape(-1, C), x.double().reshape(-1, C)
    if relu:
        g = g * ((xd * st.scale + st.shift) > 0)
    xhat = (xd - st.mean) * st.invstd
    dbeta, dgamma = g.sum(0), (g * xhat).sum(0)
    dx = st.scale / st.rows * (st.rows * g - dbeta - xhat * dgamma)
    return dx.reshape(x.shape).to(BF), dgamma.float(), dbeta.float()


def fake_relu_bwd(dy, y):
    assert dy.shape == y.shape and dy.is_contiguous() and y.is_contiguous() and dy.dtype == BF
    return (dy.double() * (y.double() > 0)).to(BF)


def fake_maxpool_fwd(x):
    pool = xb.MaxPool()
    return nhwc_bf16(pool.fwd(nchw(x))), pool


def fake_maxpool_bwd(pool, dy, in_hw):
    dx = pool.bwd(nchw(dy))
    assert tuple(dx.shape[2:]) == tuple(in_hw)
    return nhwc_bf16(dx)


def fake_nchw_to_nhwc(x, pitch=None):
    N, C, H, W = x.shape
    out = torch.zeros((N, H, W, pitch or C), dtype=BF)
    out[..., :C] = x.permute(0, 2, 3, 1).to(BF)
    return out


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def test_twin_composition_equals_blueprint(monkeypatch):
    for mod in (ops, conv_mod):
        monkeypatch.setattr(mod, "pack_conv_weight", fake_pack_conv_weight)
        monkeypatch.setattr(mod, "conv2d_nhwc", fake_conv2d_nhwc)
    monkeypatch.setattr(ops, "conv2d_wgrad", fake_conv2d_wgrad)
    monkeypatch.setattr(ops, "depthwise3x3", fake_depthwise3x3)
    monkeypatch.setattr(ops, "affine_relu", fake_affine_relu)
    monkeypatch.setattr(xt, "depthwise3x3_wgrad", fake_depthwise3x3_wgrad)
    for name, fn in (("bn_train", fake_bn_train), ("bn_relu_bwd", fake_bn_relu_bwd), ("relu_bwd", fake_relu_bwd),
                     ("maxpool3x3s2_fwd_train", fake_maxpool_fwd), ("maxpool3x3s2_bwd", fake_maxpool_bwd),
                     ("nchw_f32_to_nhwc_bf16", fake_nchw_to_nhwc)):
        monkeypatch.setattr(T, name, fn)
    assert ops.conv2d_dgrad is conv_mod.conv2d_dgrad and ops.pack_dgrad_weight is conv_mod.pack_dgrad_weight  # real

    meta = json.loads(str(np.load(GOLD)["xc_meta"]))
    scope = meta["scope"] + "/"
    heads = tuple(scope + h for h in ("rpn_head", "large_sep_feature", "final_head"))
    body = [(n[len(scope):], tuple(s)) for n, s in meta["variables"] if n.startswith(scope) and not n.startswith(heads)]
    sd = {n: torch.from_numpy(onet.seeded_variable(scope + n, s)) for n, s in body}
    rs = np.random.RandomState(5)
    images = torch.from_numpy(rs.uniform(-1, 1, (2, 3, 97, 113)).astype(np.float32))   # ragged: odd maps, uneven pads

    xb.EMULATE_BF16 = True
    try:
        tape = xb.XceptionBodyTape({k: v.double() for k, v in sd.items()})
        with torch.no_grad():
            mid0, out0 = tape.fwd(images.double())
            r_mid = torch.from_numpy(rs.standard_normal(tuple(mid0.shape))).to(BF).double()
            r_out = torch.from_numpy(rs.standard_normal(tuple(out0.shape))).to(BF).double()
            _, want = tape.bwd(r_mid, r_out)
    finally:
        xb.EMULATE_BF16 = False

    model = xt.XceptionBodyTraining(sd)
    with torch.no_grad():
        mid, out = model.fwd(images)
        grads = model.bwd(nhwc_bf16(r_mid), nhwc_bf16(r_out))
    assert mid.dtype == BF and tuple(mid.shape) == (2, mid0.shape[2], mid0.shape[3], 728)
    assert cosine(nchw(mid), mid0) > 0.99999 and cosine(nchw(out), out0) > 0.99999
    trainable = {n for n, _ in body if not n.rsplit("/", 1)[-1].startswith("moving_")}
    assert set(grads) == trainable and len(trainable) == 154
    worst = 1.0
    for n in sorted(trainable):
        assert grads[n].shape == want[n].shape, (n, grads[n].shape, want[n].shape)
        c = cosine(grads[n], want[n])
        worst = min(worst, c)
        assert c > 0.999, (n, c)
        assert abs(float(grads[n].double().norm() / want[n].norm()) - 1.0) < 0.02, n
    print("worst cosine over 154 gradients:", worst)


# ---- the registry / optimizer form (TrainableXceptionBody) over the trainer's own parameter classes -------------------
def fake_bn_train_moving(x, gamma, beta, eps, decay=None, moving_mean=None, moving_var=None):
    st = fake_bn_train(x, gamma, beta, eps)
    if moving_mean is not None:   # xdet_bn_finalize: moving <- decay*moving + (1-decay)*batch (unbiased variance)
        x2 = x.double().reshape(-1, x.shape[-1])
        moving_mean.mul_(decay).add_((1 - decay) * st.mean.float())
        moving_var.mul_(decay).add_((1 - decay) * x2.var(0, unbiased=True).float())
    return st


def fake_bn_relu_bwd_into(dy, x, st, relu, grad_view):
    dx, dgamma, dbeta = fake_bn_relu_bwd(dy, x, st, relu)
    C = x.shape[-1]
    grad_view[:C] += dbeta
    grad_view[C:2 * C] += dgamma
    return dx


def fake_sgd_momentum_conv(dw, w, mom, w_pack, w_dgrad_pack, lr, momentum, wd, grad_scale=1.0, co_off=0, ci_off=0,
                           fold=False):
    kh, kw, cin, cout = w.shape
    g = dw.view(dw.shape[0], kh, kw, dw.shape[-1])[co_off:co_off + cout, :, :, ci_off:ci_off + cin].permute(1, 2, 3, 0)
    mom.mul_(momentum).add_(g * grad_scale + wd * w)
    w.sub_(lr * mom)
    w4 = w.permute(3, 2, 0, 1).double()
    w_pack.copy_(w4)
    if w_dgrad_pack is not None:
        w_dgrad_pack.copy_(torch.flip(w4, dims=(2, 3)).permute(1, 0, 2, 3))


def fake_sgd_momentum_vec(g, w, mom, lr, momentum, wd=0.0, grad_scale=1.0):
    mom.mul_(momentum).add_(g.reshape(w.shape) * grad_scale + wd * w)
    w.sub_(lr * mom)


def test_trainable_form_accumulates_and_updates(monkeypatch):
    from xdet_b200 import light_head_rfcn_train as lt
    for mod in (ops, conv_mod):
        monkeypatch.setattr(mod, "pack_conv_weight", fake_pack_conv_weight)
        monkeypatch.setattr(mod, "conv2d_nhwc", fake_conv2d_nhwc)
    monkeypatch.setattr(ops, "conv2d_wgrad", fake_conv2d_wgrad)
    monkeypatch.setattr(ops, "depthwise3x3", fake_depthwise3x3)
    monkeypatch.setattr(ops, "affine_relu", fake_affine_relu)
    monkeypatch.setattr(xt, "depthwise3x3_wgrad", fake_depthwise3x3_wgrad)
    monkeypatch.setattr(xt, "bn_relu_bwd_into", fake_bn_relu_bwd_into)
    for name, fn in (("bn_train", fake_bn_train_moving), ("bn_relu_bwd", fake_bn_relu_bwd), ("relu_bwd", fake_relu_bwd),
                     ("maxpool3x3s2_fwd_train", fake_maxpool_fwd), ("maxpool3x3s2_bwd", fake_maxpool_bwd),
                     ("nchw_f32_to_nhwc_bf16", fake_nchw_to_nhwc), ("sgd_momentum_conv", fake_sgd_momentum_conv),
                     ("sgd_momentum_vec", fake_sgd_momentum_vec)):
        monkeypatch.setattr(T, name, fn)

    meta = json.loads(str(np.load(GOLD)["xc_meta"]))
    scope = meta["scope"] + "/"
    heads = tuple(scope + h for h in ("rpn_head", "large_sep_feature", "final_head"))
    body = [(n[len(scope):], tuple(s)) for n, s in meta["variables"] if n.startswith(scope) and not n.startswith(heads)]
    all_vars = {n: torch.from_numpy(onet.seeded_variable(scope + n, s)) for n, s in body}
    masters = {n: t.clone() for n, t in all_vars.items() if not n.rsplit("/", 1)[-1].startswith("moving_")}
    moving = {n: t.clone() for n, t in all_vars.items() if n.rsplit("/", 1)[-1].startswith("moving_")}
    rs = np.random.RandomState(9)
    images = torch.from_numpy(rs.uniform(-1, 1, (2, 3, 65, 81)).astype(np.float32))

    with torch.no_grad():
        ref = xt.XceptionBodyTraining({n: t.clone() for n, t in masters.items()})      # the standalone twin
        mid0, out0 = ref.fwd(images)
        d_mid = torch.from_numpy(rs.standard_normal(tuple(mid0.shape))).to(BF)
        d_out = torch.from_numpy(rs.standard_normal(tuple(out0.shape))).to(BF)
        want = ref.bwd(d_mid, d_out)

        reg = lt._Registry("cpu")
        net = xt.TrainableXceptionBody(masters, moving, reg, lt.ConvParams, lt.VecParam)
        reg.finalize()
        assert len(net.convs) == 4 + 2 + 34 and len(net.vecs) == 34 + 40     # convs + pointwise; depthwise + BN pairs
        mm_before = moving["block5_sepconv1_bn/moving_mean"].clone()
        mid = net.fwd_mid(images)
        out = net.fwd_exit()
        assert torch.equal(mid, mid0) and torch.equal(out, out0)
        assert net.bwd(d_mid, d_out) == {}                                    # everything went into the flat buffer
        # ---- gradients in the flat all-reduce buffer == the standalone twin's dict ----
        for layer in net._layers():
            if isinstance(layer, xt.Conv):
                leaf = "pointwise_kernel" if layer.name + "/pointwise_kernel" in masters else "kernel"
                g = layer.p.dw.view(layer.cout, layer.kh, layer.kw, -1)[..., :layer.cin].permute(1, 2, 3, 0)
                assert torch.allclose(g, want[layer.name + "/" + leaf], rtol=1e-6, atol=1e-7), layer.name
            elif isinstance(layer, xt.Depthwise):
                g = layer.vec.grad[:9 * layer.C].view(3, 3, layer.C, 1)
                assert torch.allclose(g, want[layer.name + "/depthwise_kernel"], rtol=1e-6, atol=1e-7), layer.name
            else:
                C = layer.beta.numel()
                assert torch.allclose(layer.vec.grad[:C], want[layer.name + "/beta"], rtol=1e-6, atol=1e-7)
                assert torch.allclose(layer.vec.grad[C:2 * C], want[layer.name + "/gamma"], rtol=1e-6, atol=1e-7)
        assert float(reg.flat.abs().sum()) > 0
        # ---- moving statistics followed the batch ----
        assert not torch.equal(moving["block5_sepconv1_bn/moving_mean"], mm_before)
        # ---- one optimizer step: L2 on kernels, none on the batch-norm pairs; packs and views follow ----
        before = {n: t.clone() for n, t in masters.items()}
        lr, mom, wd = 0.1, 0.9, 1e-2
        net.update(lr, mom, wd, 1.0)
        for n in ("block1_conv1/kernel", "block7_sepconv2/pointwise_kernel", "block7_sepconv2/depthwise_kernel",
                  "conv2d_4/kernel"):
            assert torch.allclose(masters[n], before[n] - lr * (want[n].float() + wd * before[n]), rtol=1e-5, atol=1e-7), n
        for n in ("block7_sepconv2_bn/gamma", "batch_normalization_2/beta"):
            assert torch.allclose(masters[n], before[n] - lr * want[n].float(), rtol=1e-5, atol=1e-7), n
        reg.flat.zero_()
        mid1, out1 = net.fwd(images)
        ref1 = xt.XceptionBodyTraining({n: t.clone() for n, t in masters.items()})   # rebuilt from the updated masters
        mid2, out2 = ref1.fwd(images)
        assert torch.equal(mid1, mid2) and torch.equal(out1, out2) and not torch.equal(out1, out0)
