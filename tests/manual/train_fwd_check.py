#!/usr/bin/env python
"""Layer-by-layer forward comparison (training-mode batch norm) GPU vs CPU oracle (debug aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import xdet_b200  # noqa: E402,F401
from oracle import net as onet  # noqa: E402
from xdet_b200 import light_head_rfcn_train as lt  # noqa: E402
from xdet_b200 import ops  # noqa: E402

size, N = 160, 2
params = lt.make_params(train_image_size=size, batch_size=N, rpn_pre_nms_top_n=600, rpn_post_nms_top_n=200,
                        rpn_min_size=16.0 / size, rpn_anchors_per_image=64, roi_one_image=32, ohem_roi_one_image=16)
tr = lt.LightHeadTrainer(params, seed=7)
sd0 = {k: v.detach().clone() for k, v in tr.store.state_dict().items()}
images, gt, gl, keys = lt.synthetic_batch(params, N, seed=3)


def rel(a, b):
    a = a.float().cpu().numpy() if torch.is_tensor(a) else a
    b = b.detach().numpy() if torch.is_tensor(b) else b
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12)), float(np.sqrt(((a - b) ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-12))


# GPU forward, manually
Wimg = size
Wo = Ho = (Wimg + 6 - 7) // 2 + 1
wp = (max((Wo - 1) * 2 + 8, Wimg + 3) + 7) // 8 * 8
x8 = ops.image_to_nhwc8(images.contiguous(), 3, wp)
y0 = ops.conv2d_nhwc(x8, tr.stem.p.pack, 64, 7, 7, padding=(3, 3, Ho, Wo), strides=(2, 2), cin=3, fold_w=(Wimg, 3))
x = ops.maxpool3x3s2_same(y0)
# CPU forward
onet.BN_TRAINING = True
nm = onet.Names(sd0)
nm.push(params["model_scope"])
with torch.no_grad():
    xc = onet.conv2d_fixed_padding(images.cpu().float(), nm, 64, 7, 2)
    print("stem conv", rel(y0.permute(0, 3, 1, 2), xc))
    xc = onet.max_pool_3x3_s2_same(xc)
    print("maxpool", rel(x.permute(0, 3, 1, 2), xc))
    cfg = ((64, 3, 1, 1), (128, 4, 2, 1), (256, 6, 2, 1), (512, 3, 2, 2))
    for li, layer in enumerate(tr.layers):
        filters, blocks, stride, dil = cfg[li]
        for bi, blk in enumerate(layer):
            x = blk.fwd(x)
            xc = onet.bottleneck_block(xc, nm, filters, bi == 0, stride if bi == 0 else 1, dilation=dil)
            print("layer%d block%d" % (li + 1, bi), rel(x.permute(0, 3, 1, 2), xc))
        if li == 2:
            rf = tr.bn_rpn.fwd(x)
            rfc = onet.batch_norm_relu(xc, nm)
            print("rpn_feat", rel(rf.permute(0, 3, 1, 2), rfc))
onet.BN_TRAINING = False
