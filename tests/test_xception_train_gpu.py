"""GPU: training-mode XceptionBody on the CUDA kernels
(x-detector_b200/net/xception_train.py) against its CPU blueprint (oracle/xception_backward.py, itself equal to
autograd): forward features and every one of the 154 gradients, on the same name-seeded variables.

Calibration done on the CPU beforehand (blueprint with EMULATE_BF16 vs the float64 blueprint, same inputs): storing
activations and activation gradients in bf16 between kernels ALONE moves the gradients of the early layers to a
cosine of 0.84 against exact arithmetic (0.93 in the exit flow; batch statistics over 2x8x8 values amplify every
rounding), so the device is compared with the bf16-EMULATING blueprint -- same rounding points, differences left are
accumulation order and the occasional flipped bf16 tie -- and only loosely with the exact one."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import net as onet
from oracle import xception_backward as xb

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "netgraph_golden.npz")


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


def test_xception_training_backbone_matches_blueprint():
    assert torch.cuda.is_available()
    import xdet_b200  # noqa: F401
    from xdet_b200 import _native
    from xdet_b200.net import xception_train as xt
    _native.lib()
    meta = json.loads(str(np.load(GOLD)["xc_meta"]))
    scope = meta["scope"] + "/"
    heads = tuple(scope + h for h in ("rpn_head", "large_sep_feature", "final_head"))
    body = [(n[len(scope):], tuple(s)) for n, s in meta["variables"] if n.startswith(scope) and not n.startswith(heads)]
    sd = {n: torch.from_numpy(onet.seeded_variable(scope + n, s)) for n, s in body}
    rs = np.random.RandomState(5)
    images = torch.from_numpy(rs.uniform(-1, 1, (2, 3, 129, 129)).astype(np.float32))
    # ---- blueprint on the CPU: exact, and with bf16 storage between kernels ----
    def blueprint(emulate, r=None):
        xb.EMULATE_BF16 = emulate
        try:
            tape = xb.XceptionBodyTape({k: v.double() for k, v in sd.items()})
            with torch.no_grad():
                m, o = tape.fwd(images.double())
                if r is None:
                    r = (torch.from_numpy(rs.standard_normal(tuple(m.shape))).bfloat16().double(),
                         torch.from_numpy(rs.standard_normal(tuple(o.shape))).bfloat16().double())
                _, g = tape.bwd(*r)
        finally:
            xb.EMULATE_BF16 = False
        return m, o, g, r
    mid_x, out_x, exact, (r_mid, r_out) = blueprint(False)
    mid0, out0, want, _ = blueprint(True, (r_mid, r_out))
    # ---- CUDA ----
    model = xt.XceptionBodyTraining({k: v.cuda() for k, v in sd.items()})
    mid, out = model.fwd(images.cuda())
    to_dev = lambda t: t.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()   # noqa: E731
    grads = model.bwd(to_dev(r_mid), to_dev(r_out))
    torch.cuda.synchronize()
    c_mid, c_out = cosine(mid.float().cpu().permute(0, 3, 1, 2), mid0), cosine(out.float().cpu().permute(0, 3, 1, 2), out0)
    c_outx = cosine(out.float().cpu().permute(0, 3, 1, 2), out_x)
    print("forward cosines: mid %.6f out %.6f out-vs-exact %.6f (blueprint bf16 vs exact: mid %.6f out %.6f)" % (
        c_mid, c_out, c_outx, cosine(mid0, mid_x), cosine(out0, out_x)))
    trainable = {n for n, _ in body if not n.rsplit("/", 1)[-1].startswith("moving_")}
    assert set(grads) == trainable and len(trainable) == 154
    rows = []
    for n in sorted(trainable):
        g = grads[n].float().cpu()
        assert g.shape == want[n].shape and torch.isfinite(g).all(), n
        rows.append((n, cosine(g, want[n]), float(g.double().norm() / (want[n].norm() + 1e-30)), cosine(g, exact[n]),
                     cosine(want[n], exact[n])))
    for r in rows:
        print("%-44s cos(bf16 blueprint) %.4f  norm ratio %.3f  cos(exact) %.4f  [blueprint bf16 vs exact %.4f]" % r)
    # Calibrated on the first B200 run (profiles/xception_train_blueprint_r2.txt): the two bf16 realisations (device, CPU
    # emulation) round at slightly different points, and batch statistics over 2x8x8 values amplify every rounding, so
    # they sit as far from each other as each sits from exact arithmetic (forward 0.998 / 0.999; early-layer gradients
    # 0.75-0.85 all three ways).  What is asserted: the device is never further from the EXACT gradient than the CPU's
    # own bf16 emulation is (margin 0.1; measured worst 0.041, and two runs of the device differ among themselves: the
    # column sums use fp32 atomics), same norms, tight in the exit flow.
    assert c_mid > 0.995, c_mid
    assert c_out > 0.99, c_out
    assert c_outx > 0.99, c_outx
    for n, c, ratio, cx, cbx in rows:
        assert cx > cbx - 0.1, (n, cx, cbx)
        assert c > (0.85 if n.startswith(("block14", "block13", "conv2d_4", "batch_normalization_4")) else 0.6), (n, c)
        assert 0.7 < ratio < 1.4, (n, ratio)
