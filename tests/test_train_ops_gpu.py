"""GPU: training-step kernels (csrc/train_ops.cu) vs the numpy oracle (target assignment, sampling: exact) and
vs plain PyTorch fp32 references of the same op (batch-norm fwd/bwd, pooling backward, losses, momentum step)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import proposals as op
from oracle import train as ot

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    assert torch.cuda.is_available()
    import xdet_b200  # noqa: F401
    from xdet_b200.ops import train
    return train


def make_gt(rng, N, G):
    gt = np.zeros((N, G, 4), np.float32)
    gl = np.zeros((N, G), np.int32)
    for n in range(N):
        k = rng.integers(1, G + 1)
        cy, cx = rng.uniform(0.2, 0.8, k), rng.uniform(0.2, 0.8, k)
        h, w = rng.uniform(0.1, 0.5, k), rng.uniform(0.1, 0.5, k)
        gt[n, :k] = np.clip(np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], -1), 0, 1)
        gl[n, :k] = rng.integers(1, 21, k)
    return gt, gl


def test_anchor_match_encode_exact(T):
    rng = np.random.default_rng(0)
    y, x, h, w = op.layer_anchors((480, 480), (30, 30), [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5], 16)
    cy = np.broadcast_to(y[:, :, None], (30, 30, 22)).reshape(-1).astype(np.float32)
    cx = np.broadcast_to(x[:, :, None], (30, 30, 22)).reshape(-1).astype(np.float32)
    hh = np.broadcast_to(h[None, None, :], (30, 30, 22)).reshape(-1).astype(np.float32)
    ww = np.broadcast_to(w[None, None, :], (30, 30, 22)).reshape(-1).astype(np.float32)
    ref = np.stack([cy, cx, hh, ww], -1)
    pts = np.stack([cy - hh / np.float32(2), cx - ww / np.float32(2), cy + hh / np.float32(2), cx + ww / np.float32(2)], -1)
    gt, gl = make_gt(rng, 4, 6)
    lab, tgt, sc = T.match_encode(torch.from_numpy(pts).cuda(), torch.from_numpy(gt).cuda(), torch.from_numpy(gl).cuda(),
                                  0.0, 0.7, 0.3, ref_yxhw=torch.from_numpy(ref).cuda())
    for n in range(4):
        l0, t0, s0 = ot.match_encode(pts, gt[n], gl[n], 0.0, 0.7, 0.3, ref_yxhw=ref)
        assert np.array_equal(lab[n].cpu().numpy(), l0)
        assert np.array_equal(sc[n].cpu().numpy().view(np.int32), s0.view(np.int32))
        assert np.abs(tgt[n].cpu().numpy() - t0).max() < 2e-6
        assert (l0 > 0).sum() >= 1


def test_roi_match_encode_and_sampling_exact(T):
    rng = np.random.default_rng(1)
    N, R, G = 3, 1800, 6
    gt, gl = make_gt(rng, N, G)
    cy, cx = rng.uniform(0.1, 0.9, (N, R)), rng.uniform(0.1, 0.9, (N, R))
    h, w = rng.uniform(0.05, 0.6, (N, R)), rng.uniform(0.05, 0.6, (N, R))
    rois = np.clip(np.stack([cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2], -1), 0, 1).astype(np.float32)
    rois[:, :G] = gt + rng.normal(0, 0.01, gt.shape).astype(np.float32) * (gl[..., None] > 0)  # some strong overlaps
    rois = np.clip(rois, 0, 1).astype(np.float32)
    lab, tgt, sc = T.match_encode(torch.from_numpy(rois).cuda(), torch.from_numpy(gt).cuda(), torch.from_numpy(gl).cuda(),
                                  1.0, 0.53, 0.5)
    keys = [rng.random((N, R), dtype=np.float32) for _ in range(2)] + [rng.random((N, 64), dtype=np.float32)]
    idx, cnt = T.sample_fg_bg(lab, sc, 0.0, 16, 64, *[torch.from_numpy(k).cuda() for k in keys])
    for n in range(N):
        l0, t0, s0 = ot.match_encode(rois[n], gt[n], gl[n], 1.0, 0.53, 0.5)
        assert np.array_equal(lab[n].cpu().numpy(), l0)
        assert np.array_equal(sc[n].cpu().numpy().view(np.int32), s0.view(np.int32))
        ok = np.isfinite(t0).all(axis=1)
        assert np.abs(tgt[n].cpu().numpy()[ok] - t0[ok]).max() < 2e-6
        i0, c0 = ot.sample_fg_bg(l0, s0, 0.0, 16, 64, keys[0][n], keys[1][n], keys[2][n])
        assert tuple(cnt[n].cpu().numpy()) == c0
        assert np.array_equal(idx[n].cpu().numpy(), i0)


@pytest.mark.parametrize("case", [(5000, 40, 128, 256), (158400, 3000, 1024, 2048), (200, 3, 16, 64), (300, 0, 16, 64)])
def test_sampling_regimes(T, case):
    """down-sampling of both classes, up-sampling when candidates are short, no positives at all."""
    n, npos, exp_fg, total = case
    rng = np.random.default_rng(n)
    labels = np.zeros(n, np.int32)
    labels[rng.choice(n, npos, replace=False)] = 1
    labels[rng.choice(n, n // 10, replace=False)] -= 1  # some ignore (-1) / some positives turned background
    if n == 200:
        labels[labels == 0] = -1
        labels[:20] = 0  # 20 negatives only: up-sampling
    keys = [rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32), rng.random(total, dtype=np.float32)]
    idx, cnt = T.sample_fg_bg(torch.from_numpy(labels[None]).cuda(), None, 0.0, exp_fg, total,
                              *[torch.from_numpy(k[None]).cuda() for k in keys])
    i0, c0 = ot.sample_fg_bg(labels, None, 0.0, exp_fg, total, *keys)
    assert tuple(cnt[0].cpu().numpy()) == c0
    assert np.array_equal(idx[0].cpu().numpy(), i0)


def test_bn_train_forward_backward(T):
    import xdet_b200.ops as ops
    g = torch.Generator(device="cuda").manual_seed(0)
    N, H, W, C = 4, 30, 30, 264
    x = (torch.randn((N, H, W, C), generator=g, device="cuda") * 1.5 + 0.3).to(torch.bfloat16)
    gamma = torch.rand(C, generator=g, device="cuda") + 0.5
    beta = torch.randn(C, generator=g, device="cuda") * 0.2
    mm, mv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    st = T.bn_train(x, gamma, beta, 1e-5, 0.997, mm, mv)
    y = ops.affine_relu(x, st.scale, st.shift, relu=True)
    xf = x.float().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    yr = torch.relu(F.batch_norm(xf.reshape(-1, C), rm, rv, gr, br, training=True, momentum=1 - 0.997, eps=1e-5))
    assert (y.float().reshape(-1, C) - yr).abs().max().item() < 0.03
    assert (mm - rm).abs().max().item() < 1e-5 and (mv - rv).abs().max().item() < 1e-5
    dy = torch.randn((N, H, W, C), generator=g, device="cuda").to(torch.bfloat16)
    add = torch.randn((N, H, W, C), generator=g, device="cuda").to(torch.bfloat16)
    yr.backward(dy.float().reshape(-1, C))
    dx, dgamma, dbeta = T.bn_relu_bwd(dy, x, st, relu=True, add_in=add)
    torch.cuda.synchronize()
    ref_dx = xf.grad + add.float()
    assert (dx.float() - ref_dx).abs().max().item() < 0.03 * max(1.0, ref_dx.abs().max().item())
    assert (dgamma - gr.grad).abs().max().item() < 2e-2 * gr.grad.abs().max().item()
    assert (dbeta - br.grad).abs().max().item() < 2e-2 * br.grad.abs().max().item()


@pytest.mark.parametrize("rows,C", [(7200, 256), (7200, 2048), (115200, 64), (28800, 496), (37, 8), (5000, 728)])
def test_bn_train_stats_one_launch_over_a_self_cleaning_scratch(T, rows, C):
    """xdet_bn_train_stats_bf16 = column statistics + bn_finalize in one launch: against float64 statistics of the same
    bf16 values; called back to back with different widths on the same scratch (the kernel must hand it back zeroed);
    the backward's reduce with the raw sum of g*x against the direct sum of g*xhat."""
    g = torch.Generator(device="cuda").manual_seed(rows + C)
    for rep in range(3):   # same stream, same scratch, three layers in a row
        x = (torch.randn((rows, C), generator=g, device="cuda") * (1.0 + rep) + 0.7 * rep).to(torch.bfloat16)
        gamma = torch.rand(C, generator=g, device="cuda") + 0.5
        beta = torch.randn(C, generator=g, device="cuda") * 0.2
        mm, mv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        st = T.bn_train(x, gamma, beta, 1e-5, 0.997, mm, mv)
        xd = x.double()
        mean, var = xd.mean(0), xd.var(0, unbiased=False)
        inv = 1.0 / torch.sqrt(var + 1e-5)
        assert (st.mean.double() - mean).abs().max().item() < 2e-5 * (1 + mean.abs().max().item())
        assert ((st.invstd.double() - inv).abs() / inv).max().item() < 1e-4
        assert ((st.scale.double() - gamma.double() * inv).abs() / inv).max().item() < 1e-4
        assert (st.shift.double() - (beta.double() - mean * gamma.double() * inv)).abs().max().item() < 2e-4
        unb = var * rows / max(rows - 1, 1)
        assert (mm.double() - 0.003 * mean).abs().max().item() < 1e-6
        assert (mv.double() - (0.997 + 0.003 * unb)).abs().max().item() < 1e-5 * (1 + unb.max().item())
        assert int(T._bn_scratch(x.device, C).view(torch.int32).abs().sum()) == 0
        dy = torch.randn((rows, C), generator=g, device="cuda").to(torch.bfloat16)
        pre = xd * st.scale.double() + st.shift.double()
        dy[pre.abs() < 1e-5 * ((xd * st.scale.double()).abs() + st.shift.double().abs())] = 0   # mask a rounding away
        dx, dgamma, dbeta = T.bn_relu_bwd(dy, x, st, relu=True)
        mask = (pre > 0).double()
        gd = dy.double() * mask
        xhat = (xd - st.mean.double()) * st.invstd.double()
        ref_b, ref_g = gd.sum(0), (gd * xhat).sum(0)
        tol = 1e-4 * (gd.abs() * (1 + xhat.abs())).sum(0).max().item()
        assert (dbeta.double() - ref_b).abs().max().item() < tol and (dgamma.double() - ref_g).abs().max().item() < tol
        ref_dx = st.scale.double() * (gd - ref_b / rows - xhat * ref_g / rows)
        assert (dx.double() - ref_dx).abs().max().item() < 0.01 * max(1.0, ref_dx.abs().max().item())


@pytest.mark.parametrize("case", [
    # N, H, W, cin, cout, k, stride, dil, residual
    (8, 30, 30, 256, 1024, 1, 1, 1, True),     # flattened 1x1 with the residual epilogue, 4 N tiles of 256
    (2, 120, 120, 64, 64, 3, 1, 1, False),     # wide rows: 120 of 128 tile columns are inside the image
    (3, 61, 45, 128, 136, 3, 2, 1, False),     # ragged tiles in both directions, a partial channel chunk, stride 2
    (8, 30, 30, 512, 512, 3, 1, 2, False),     # the dilated block_layer4 shape
    (4, 30, 30, 2048, 512, 1, 1, 1, False),
])
def test_conv_epilogue_batch_statistics(T, case):
    """conv2d_nhwc(stats=...): the epilogue's per-channel sums / sums of squares of the STORED bf16 output (rows outside
    the image and channels outside Cout excluded, statistics added to what the buffer holds), then bn_train_apply against
    the standalone statistics kernel + affine_relu on the same tensor."""
    import xdet_b200.ops as ops
    N, H, W, cin, cout, k, stride, dil, with_res = case
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    x = torch.randn((N, H, W, cin), generator=g, device="cuda").to(torch.bfloat16)
    w = torch.randn((cout, cin, k, k), generator=g, device="cuda") * (1.0 / (cin * k * k) ** 0.5)
    wp = ops.pack_conv_weight(w)
    bias = torch.randn(cout, generator=g, device="cuda") * 0.3
    Ho, Wo = -(-H // stride), -(-W // stride)
    res = torch.randn((N, Ho, Wo, cout), generator=g, device="cuda").to(torch.bfloat16) if with_res else None
    pre = torch.full((2 * cout,), 0.5, device="cuda")     # the kernel ADDS
    sums = pre.clone()
    y = ops.conv2d_nhwc(x, wp, cout, k, k, dilation=(dil, dil), strides=(stride, stride), bias=bias, residual=res,
                        stats=sums)
    y_plain = ops.conv2d_nhwc(x, wp, cout, k, k, dilation=(dil, dil), strides=(stride, stride), bias=bias, residual=res)
    torch.cuda.synchronize()
    assert torch.equal(y, y_plain)
    yd = y.double().reshape(-1, cout)
    ref = torch.cat([yd.sum(0), (yd * yd).sum(0)]) + 0.5
    scale = torch.cat([yd.abs().sum(0), (yd * yd).sum(0)]) + 1.0
    assert ((sums.double() - ref).abs() / scale).max().item() < 2e-6
    gamma = torch.rand(cout, generator=g, device="cuda") + 0.5
    beta = torch.randn(cout, generator=g, device="cuda") * 0.2
    mm, mv = torch.zeros(cout, device="cuda"), torch.ones(cout, device="cuda")
    mm2, mv2 = mm.clone(), mv.clone()
    z, st = T.bn_train_apply(y, sums - pre, gamma, beta, 1e-5, 0.997, mm, mv, relu=True)
    st0 = T.bn_train(y, gamma, beta, 1e-5, 0.997, mm2, mv2)
    z0 = ops.affine_relu(y, st0.scale, st0.shift, relu=True)
    torch.cuda.synchronize()
    assert st.rows == st0.rows
    for a, b in ((st.mean, st0.mean), (st.invstd, st0.invstd), (st.scale, st0.scale), (st.shift, st0.shift), (mm, mm2),
                 (mv, mv2)):
        assert ((a - b).abs() / (1e-3 + b.abs())).max().item() < 2e-4
    assert (z.float() - z0.float()).abs().max().item() <= 0.02 * max(1.0, z0.float().abs().max().item())
    # and bit for bit when fed the very same sums
    sums2 = T.col_stats(y.reshape(-1, cout), True)
    z2, st2 = T.bn_train_apply(y, sums2, gamma, beta, 1e-5)
    assert torch.equal(z2, ops.affine_relu(y, st2.scale, st2.shift, relu=True))


def test_maxpool_backward(T):
    import xdet_b200.ops as ops
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn((2, 60, 60, 64), generator=g, device="cuda").to(torch.bfloat16)
    y, arg = T.maxpool3x3s2_fwd_train(x)
    assert torch.equal(y, ops.maxpool3x3s2_same(x))
    dy = torch.randn(y.shape, generator=g, device="cuda").to(torch.bfloat16)
    dx = T.maxpool3x3s2_bwd(arg, dy, x.shape[1:3])
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    yr = F.max_pool2d(F.pad(xr, (0, 1, 0, 1), value=float("-inf")), 3, 2)  # 60 -> 30: SAME pads (0,1)
    assert torch.equal(yr.permute(0, 2, 3, 1), y.float())
    yr.backward(dy.float().permute(0, 3, 1, 2))
    # ties between equal bf16 values may be routed differently; everything else is exact up to bf16 rounding of sums
    diff = (dx.float() - xr.grad.permute(0, 2, 3, 1)).abs()
    assert (diff > 0.02).float().mean().item() < 2e-3


def test_losses(T):
    g = torch.Generator(device="cuda").manual_seed(3)
    M = 777
    logits = torch.randn((M, 25), generator=g, device="cuda") * 3
    labels = torch.randint(0, 21, (M,), generator=g, device="cuda", dtype=torch.int32)
    rw = torch.rand(M, generator=g, device="cuda")
    loss, dl = T.softmax_ce(logits, labels, 21, row_w=rw, w_all=0.5)
    lr_ = logits[:, :21].clone().requires_grad_(True)
    ref = F.cross_entropy(lr_, labels.long(), reduction="none")
    (ref * rw * 0.5).sum().backward()
    assert (loss - ref).abs().max().item() < 1e-5
    assert (dl[:, :21] - lr_.grad).abs().max().item() < 1e-6 and dl[:, 21:].abs().max().item() == 0
    pred = (torch.randn((M, 25), generator=g, device="cuda") * 2)
    tgt = torch.randn((M, 4), generator=g, device="cuda")
    l1, dp = T.smooth_l1(pred[:, 21:], tgt, row_w=rw, w_all=2.0)
    pr = pred[:, 21:].clone().requires_grad_(True)
    ref1 = F.smooth_l1_loss(pr, tgt, reduction="none", beta=1.0).sum(-1) * rw
    (ref1 * 2.0).sum().backward()
    assert (l1 - ref1).abs().max().item() < 1e-5 and (dp - pr.grad).abs().max().item() < 1e-6


def test_sgd_momentum_and_repack(T):
    import xdet_b200.ops as ops
    g = torch.Generator(device="cuda").manual_seed(4)
    kh, kw, cin, cout = 3, 3, 72, 40
    w = torch.randn((kh, kw, cin, cout), generator=g, device="cuda") * 0.1
    mom = torch.randn(w.shape, generator=g, device="cuda") * 0.01
    dw_tf = torch.randn(w.shape, generator=g, device="cuda")
    cpad = 128
    dw = torch.zeros((cout, kh * kw, cpad), device="cuda")
    dw[:, :, :cin] = dw_tf.permute(3, 0, 1, 2).reshape(cout, kh * kw, cin)
    wp = torch.zeros((cout, kh * kw * cpad), dtype=torch.bfloat16, device="cuda")
    wd_pack = torch.zeros((cin, kh * kw * 64), dtype=torch.bfloat16, device="cuda")
    w0, m0 = w.clone(), mom.clone()
    T.sgd_momentum_conv(dw, w, mom, wp, wd_pack, 0.01, 0.9, 2e-4)
    a = 0.9 * m0 + dw_tf + 2e-4 * w0
    assert (mom - a).abs().max().item() < 1e-6 and (w - (w0 - 0.01 * a)).abs().max().item() < 1e-6
    assert torch.equal(wp, ops.pack_conv_weight(w.permute(3, 2, 0, 1)))
    assert torch.equal(wd_pack, ops.pack_dgrad_weight(w.permute(3, 2, 0, 1)))
    v, mv, gv = torch.randn(100, device="cuda"), torch.zeros(100, device="cuda"), torch.randn(100, device="cuda")
    v0 = v.clone()
    T.sgd_momentum_vec(gv, v, mv, 0.1, 0.9)
    assert (v - (v0 - 0.1 * gv)).abs().max().item() < 1e-6


def test_sgd_multi_matches_the_per_variable_kernels(T):
    """xdet_sgd_momentum_multi: several convolutions (one a fused pack of two masters at channel offsets, one dense,
    one without an input-gradient pack) and vectors with different decay in ONE launch = the per-variable launches,
    bit for bit."""
    import xdet_b200.ops as ops
    g = torch.Generator(device="cuda").manual_seed(14)
    rnd = lambda *s: torch.randn(s, generator=g, device="cuda")  # noqa: E731

    def conv_case(kh, kw, cin, couts, need_d=True):
        cout = sum(couts)
        cpad, copad = (cin + 63) // 64 * 64, (cout + 63) // 64 * 64
        masters, co = [], 0
        for c in couts:
            shape = (cin, c) if kh == kw == 1 and len(couts) == 1 and not need_d else (kh, kw, cin, c)
            masters.append((rnd(*shape) * 0.1, rnd(*shape) * 0.01, co))
            co += c
        dw = rnd(cout, kh * kw, cpad)
        wp = torch.zeros((cout, kh * kw * cpad), dtype=torch.bfloat16, device="cuda")
        dp = torch.zeros((cin, kh * kw * copad), dtype=torch.bfloat16, device="cuda") if need_d else None
        return masters, dw, wp, dp

    cases = [conv_case(3, 3, 72, [40]), conv_case(1, 1, 256, [64, 100]), conv_case(1, 1, 130, [50], need_d=False),
             conv_case(15, 1, 33, [24])]
    vecs = [(rnd(100), rnd(100), rnd(104), 2e-4), (rnd(2048), rnd(2048), rnd(2048), 0.0), (rnd(7), rnd(7), rnd(8), 0.0)]
    clone = lambda t: None if t is None else t.clone()  # noqa: E731
    ref_cases = [([tuple(clone(x) if torch.is_tensor(x) else x for x in m) for m in ms], dw, clone(wp), clone(dp))
                 for ms, dw, wp, dp in cases]
    ref_vecs = [(w.clone(), m.clone(), gv, wd) for w, m, gv, wd in vecs]
    lr, mo, gs = 0.013, 0.9, 0.5
    for ms, dw, wp, dp in ref_cases:
        for w, m, co in ms:
            T.sgd_momentum_conv(dw, w, m, wp, dp, lr, mo, 1e-4, gs, co, 0)
    for w, m, gv, wd in ref_vecs:
        T.sgd_momentum_vec(gv[:w.numel()], w, m, lr, mo, wd, gs)
    plan = T.SgdPlan()
    for ms, dw, wp, dp in cases:
        for w, m, co in ms:
            plan.add_conv(dw, w, m, wp, dp, 1e-4, co, 0)
    for w, m, gv, wd in vecs:
        plan.add_vec(gv[:w.numel()], w, m, wd)
    plan.step(lr, mo, gs)
    torch.cuda.synchronize()
    for (ms, _, wp, dp), (rms, _, rwp, rdp) in zip(cases, ref_cases):
        for (w, m, _), (rw, rm, _) in zip(ms, rms):
            assert torch.equal(w, rw) and torch.equal(m, rm)
        assert torch.equal(wp, rwp) and (dp is None or torch.equal(dp, rdp))
        w4 = torch.cat([(w if w.dim() == 4 else w.reshape(1, 1, *w.shape)) for w, _, _ in ms], dim=3)
        assert torch.equal(wp, ops.pack_conv_weight(w4.permute(3, 2, 0, 1)))
    for (w, m, _, _), (rw, rm, _, _) in zip(vecs, ref_vecs):
        assert torch.equal(w, rw) and torch.equal(m, rm)


def test_thin_map_repacks(T):
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((2, 490, 30, 30), generator=g, device="cuda")
    y = T.nchw_f32_to_nhwc_bf16(x)
    assert torch.equal(y, x.permute(0, 2, 3, 1).to(torch.bfloat16))
    sc, sh = torch.rand(490, device="cuda") + 0.5, torch.randn(490, device="cuda")
    z = T.affine_relu_to_nchw_f32(y, sc, sh, relu=True)
    ref = torch.relu(y.float() * sc + sh).permute(0, 3, 1, 2)
    assert (z - ref).abs().max().item() < 1e-5
