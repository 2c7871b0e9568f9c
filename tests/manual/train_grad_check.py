#!/usr/bin/env python
"""Per-variable comparison of the GPU training step's gradients with the CPU autograd oracle (debug aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import xdet_b200  # noqa: E402,F401
from oracle import net_train as ont  # noqa: E402
from oracle import proposals as op  # noqa: E402
from xdet_b200 import light_head_rfcn_train as lt  # noqa: E402

size = int(os.environ.get("SIZE", "160"))
N = int(os.environ.get("BATCH", "2"))
params = lt.make_params(train_image_size=size, batch_size=N, rpn_pre_nms_top_n=600, rpn_post_nms_top_n=200,
                        rpn_min_size=16.0 / size, rpn_anchors_per_image=64, roi_one_image=32, ohem_roi_one_image=16,
                        resnet_layers=tuple(int(v) for v in os.environ.get("LAYERS", "3,4,6,3").split(",")))
tr = lt.LightHeadTrainer(params, seed=7)
sd0 = {k: v.detach().clone() for k, v in tr.store.state_dict().items()}
batch = lt.synthetic_batch(params, N, seed=3)
out = tr.step(*batch, apply_update=False)
torch.cuda.synchronize()
images, gt, gl, keys = batch
fm = size // 16
anchors = op.layer_anchors((size, size), (fm, fm), [0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8], [0.1], [1., 2., .5], 16)
inject = {k: out[k].cpu().numpy() for k in ("rpn_idx", "rois_all", "roi_idx", "ohem_idx")}
losses, grads, mid = ont.train_step(images.cpu().numpy(), gt.cpu().numpy(), gl.cpu().numpy(), sd0, params, anchors, inject)
print("losses gpu", {k: float(out[k]) for k in losses}, "cpu", losses)
t = out["large_sep_feature"].cpu().numpy()
print("thin rel err", np.abs(t - mid["large_sep_feature"]).max() / np.abs(mid["large_sep_feature"]).max())
for cp in tr.convs:
    for i, (key, tns, co, ci) in enumerate(cp.masters):
        kh, kw = cp.kh, cp.kw
        cin, cout = (tns.shape if tns.dim() == 2 else (tns.shape[2], tns.shape[3]))
        if cp.fold:
            d = cp.dw.reshape(cp.cout, kh, 8, 8)[:, :, :kw, :cin].permute(1, 2, 3, 0)
        else:
            d = cp.dw[co:co + cout, :, ci:ci + cin].reshape(cout, kh, kw, cin).permute(1, 2, 3, 0)
        a, b = d.float().cpu().flatten(), grads[key].flatten()
        cos = float(torch.dot(a, b) / (a.norm() * b.norm() + 1e-20))
        print("%-70s cos %.4f  |gpu| %.4e |cpu| %.4e" % (key, cos, float(a.norm()), float(b.norm())))
