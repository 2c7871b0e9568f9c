"""CPU restatement of ONE training step of the reference's lighr_head_model_fn (light_head_rfcn_train.py:277-451)
for the ResNet-50 light-head composition (and, with params['backbone'] = 'xception', the reference's XceptionBody):
PyTorch-CPU fp32 graph + autograd.  TEST INFRASTRUCTURE ONLY.

The discrete selections (RPN sample indices, sampled RoIs, OHEM indices) can be INJECTED so that gradients are
compared on identical decisions (they are discontinuous in the bf16-vs-fp32 rounding); the selections themselves
are checked separately and exactly (oracle/train.py vs the kernels).

PINNED (forward) by the reference's own model_fn: tests/golden/make_trainstep_golden.py calls the unmodified
lighr_head_model_fn in TRAIN mode under the numpy TensorFlow stand-in (oracle/tf_shim; the custom op bound to the
reference's compiled PsRoIAlign) and tests/test_trainstep_golden.py holds this file to its RPN sample indices, RoI
selection, OHEM indices and the three loss values.  Gradients are torch autograd over that forward; TensorFlow's own
gradient code is not available offline, so the backward stays unpinned by the reference."""
import numpy as np
import torch

from . import net as onet
from . import proposals as P
from . import psroi
from . import train as ot


class _PsRoi(torch.autograd.Function):
    @staticmethod
    def forward(ctx, thin, yxhw):
        pooled, index = psroi.psroi_align_fwd(thin.detach().numpy(), yxhw, 7, 7, "max")
        ctx.shape, ctx.yxhw, ctx.index = tuple(thin.shape), yxhw, index
        return torch.from_numpy(pooled).to(thin.dtype)

    @staticmethod
    def backward(ctx, g):
        d = psroi.psroi_align_bwd(ctx.shape, ctx.yxhw, np.ascontiguousarray(g.numpy(), dtype=np.float32), ctx.index, 7, 7,
                                  "max")
        return torch.from_numpy(d).to(g.dtype), None


def smooth_l1(x):  # modified_smooth_l1, sigma = 1 (:257-275)
    ax = x.abs()
    return torch.where(ax < 1.0, 0.5 * x * x, ax - 0.5)


def train_step(images, gt_boxes, gt_labels, sd, params, anchors, inject, dtype=torch.float32):
    """-> (losses dict, grads {variable name: d total_loss_without_L2 / d variable}, intermediates).
    ``dtype=torch.float64``: the same graph in double precision (PsRoIAlign stays the fp32 C oracle, its output cast up)
    -- the arbiter that tells whose fp32 rounding a gradient difference is."""
    y, x, h, w = anchors
    fm, A = y.shape[0], h.shape[0]
    cy = np.broadcast_to(y[:, :, None], (fm, fm, A)).reshape(-1).astype(np.float32)
    cx = np.broadcast_to(x[:, :, None], (fm, fm, A)).reshape(-1).astype(np.float32)
    hh = np.broadcast_to(h[None, None, :], (fm, fm, A)).reshape(-1).astype(np.float32)
    ww = np.broadcast_to(w[None, None, :], (fm, fm, A)).reshape(-1).astype(np.float32)
    ref = np.stack([cy, cx, hh, ww], -1)
    pts = np.stack([cy - hh / np.float32(2), cx - ww / np.float32(2), cy + hh / np.float32(2), cx + ww / np.float32(2)], -1)
    N = images.shape[0]
    nc = params["num_classes"]
    nm = onet.Names(sd)
    nm.leaves = {}
    nm.push(params["model_scope"])
    onet.BN_TRAINING = True
    prev_dtype, onet.DTYPE = onet.DTYPE, dtype
    try:
        xin = torch.as_tensor(images).to(dtype)
        if params.get("backbone", "resnet50") == "xception":  # the reference's own backbone (net/xception_body.py:236-379)
            rpn_feat, backbone = onet.xception_body(xin, nm)
        else:
            rpn_feat, backbone = onet.lighthead_resnet50_body(xin, nm,
                                                              layers=tuple(params.get("resnet_layers", (3, 4, 6, 3))))
        cls, box = onet.get_rpn(rpn_feat, nm, "rpn_head", num_anchors=A)
        thin = onet.large_sep_kernel(backbone, nm, "large_sep_feature")
        cls_all = cls.permute(0, 2, 3, 1).reshape(-1, 2)
        loc_all = box.permute(0, 2, 3, 1).reshape(-1, 4)
        # anchor targets (the reference's input pipeline: encode_all_anchors)
        gl, gt = [], []
        for n in range(N):
            l0, t0, _ = ot.match_encode(pts, gt_boxes[n], gt_labels[n], 0.0, params["rpn_match_threshold"],
                                        params["rpn_neg_threshold"], ref_yxhw=ref)
            gl.append(l0)
            gt.append(t0)
        glabels, gtargets = np.concatenate(gl), np.concatenate(gt)
        idx = torch.as_tensor(np.asarray(inject["rpn_idx"]), dtype=torch.long)
        s_lab = torch.as_tensor((np.clip(glabels, 0, nc)[idx.numpy()] > 0).astype(np.int64))
        rpn_ce = torch.nn.functional.cross_entropy(cls_all[idx], s_lab)
        pos = s_lab > 0
        l1 = smooth_l1(loc_all[idx][pos] - torch.as_tensor(gtargets[idx.numpy()][pos.numpy()])).sum(-1)
        rpn_loc = l1.mean() / params["rpn_fg_ratio"]
        # sampled RoIs and their targets
        rois_all = np.asarray(inject["rois_all"], np.float32)
        roi_idx = np.asarray(inject["roi_idx"])
        R = roi_idx.shape[1]
        rois, rtgt, rlab = np.zeros((N, R, 4), np.float32), np.zeros((N, R, 4), np.float32), np.zeros((N, R), np.int64)
        for n in range(N):
            l0, t0, _ = ot.match_encode(rois_all[n], gt_boxes[n], gt_labels[n], 0.1, params["match_threshold"],
                                        params["neg_threshold_high"])
            rois[n], rtgt[n], rlab[n] = rois_all[n][roi_idx[n]], t0[roi_idx[n]], l0[roi_idx[n]]
        yxhw = P.point2center(rois.reshape(-1, 4)).reshape(N, R, 4).astype(np.float32)
        pooled = _PsRoi.apply(thin, np.ascontiguousarray(yxhw)).reshape(N, R, -1)
        sel = np.asarray(inject["ohem_idx"])  # [N,k]; tf.gather(..., axis=1): every image gets every row of indices
        k = sel.shape[1]
        sel_t = torch.as_tensor(sel.reshape(-1), dtype=torch.long)
        feat2 = pooled[:, sel_t].reshape(N * N * k, -1)
        lab2 = torch.as_tensor(rlab[:, sel.reshape(-1)].reshape(-1))
        tgt2 = torch.as_tensor(rtgt[:, sel.reshape(-1)].reshape(-1, 4))
        nm.push("final_head")
        hfc = onet.dense(feat2, nm, "subnet_fc", 2048, relu=True)
        cls_score = onet.dense(hfc, nm, "fc_cls", nc)
        bbox_reg = onet.dense(hfc, nm, "fc_loc", 4)
        nm.pop()
        ce = torch.nn.functional.cross_entropy(cls_score, lab2, reduction="none")
        loc = smooth_l1(bbox_reg - tgt2).sum(-1) * (lab2 > 0).float()
        head = (ce + loc / params["fg_ratio"]).mean()
        total = rpn_ce + rpn_loc + head
        total.backward()
    finally:
        onet.BN_TRAINING = False
        onet.DTYPE = prev_dtype
    grads = {k_: (v.grad if v.grad is not None else torch.zeros_like(v)) for k_, v in nm.leaves.items()}
    return ({"rpn_cross_entropy_loss": float(rpn_ce.detach()), "rpn_location_loss": float(rpn_loc.detach()),
             "head_loss": float(head.detach())}, grads,
            {"glabels": glabels, "roi_labels": rlab, "roi_targets": rtgt, "rois": rois,
             "large_sep_feature": thin.detach().numpy()})
