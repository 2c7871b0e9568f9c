"""CPU: the hand-written backward of the training-mode XceptionBody (oracle/xception_backward.py -- the op sequence
the GPU trainer will launch: wgrad / dgrad per conv, depthwise input gradient as the forward kernel with flipped taps,
batch-norm backward from batch statistics with and without the ReLU mask, max-pool scatter + residual) against torch
autograd over oracle/net.xception_body, in float64, for every one of the 154 trainable variables and the image."""
import json
import os

import numpy as np
import torch

from oracle import net as onet
from oracle import xception_backward as xb

GOLD = os.path.join(os.path.dirname(__file__), "golden", "netgraph_golden.npz")


class Names64(onet.Names):
    """oracle/net.py's variable lookup with float64 autograd leaves."""

    def get(self, *parts, shape=None):
        key = "/".join(self.scope + list(parts))
        if key not in self.leaves:
            self.leaves[key] = self.sd[key].double().clone().requires_grad_(True)
        return self.leaves[key]


def test_manual_backward_equals_autograd():
    meta = json.loads(str(np.load(GOLD)["xc_meta"]))
    scope = meta["scope"] + "/"
    body = [(n[len(scope):], tuple(s)) for n, s in meta["variables"]
            if n.startswith(scope) and "/" not in n[len(scope):].rsplit("/", 1)[0]
            and not n.startswith((scope + "rpn_head", scope + "large_sep_feature", scope + "final_head"))]
    sd = {n: torch.from_numpy(onet.seeded_variable(scope + n, s)).double() for n, s in body}
    rs = np.random.RandomState(5)
    x = torch.from_numpy(rs.uniform(-1, 1, (2, 3, 65, 65)))
    # ---- autograd over the oracle graph (batch statistics) ----
    nm = Names64(sd)
    nm.leaves = {}
    xa = x.clone().requires_grad_(True)
    prev, onet.BN_TRAINING = onet.BN_TRAINING, True
    try:
        mid, out = onet.xception_body(xa, nm)
    finally:
        onet.BN_TRAINING = prev
    r_mid = torch.from_numpy(rs.standard_normal(tuple(mid.shape)))
    r_out = torch.from_numpy(rs.standard_normal(tuple(out.shape)))
    ((mid * r_mid).sum() + (out * r_out).sum()).backward()
    # ---- the explicit tape ----
    tape = xb.XceptionBodyTape(sd)
    with torch.no_grad():
        mid2, out2 = tape.fwd(x)
        dx, grads = tape.bwd(r_mid, r_out)
    assert torch.allclose(mid2, mid.detach(), rtol=1e-10, atol=1e-10) and torch.allclose(out2, out.detach(), rtol=1e-10, atol=1e-10)
    trainable = {n for n, _ in body if not n.rsplit("/", 1)[-1].startswith("moving_")}
    assert set(grads) == trainable and len(trainable) == 154
    worst = 0.0
    for n in sorted(trainable):
        want = nm.leaves[n].grad
        assert want is not None and grads[n].shape == want.shape, n
        err = float((grads[n] - want).abs().max() / (want.abs().max() + 1e-30))
        worst = max(worst, err)
        assert err < 1e-8, (n, err)
    assert float((dx - xa.grad).abs().max() / xa.grad.abs().max()) < 1e-8
    assert worst > 0.0   # not trivially identical objects
