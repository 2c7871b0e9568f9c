"""Light-Head R-CNN inference entry point -- the surface of the reference's ``light_head_rfcn_eval.py``.

* ``FLAGS`` carries the reference's flag names and defaults (light_head_rfcn_eval.py:41-139).
* ``lighr_head_model_fn(features, labels, mode, params)`` (reference :364-445, spelling kept) runs the graph
  up to ``head_bboxes_pred`` / ``head_cls_score`` -- backbone -> get_rpn -> large_sep_kernel -> objectness /
  decode_all_anchors -> get_proposals -> get_head -> ext_decode_rois + softmax -- entirely on the GPU (the
  reference bounces through /cpu:0 for proposals).
* ``bboxes_eval`` (:263-290): the per-class select / clip / filter / sort / NMS post-processing, one batched GPU
  selection.  TP/FP matching and VOC mAP (:292-362) are SURVEY 8 f3 and are not built.
* ``main`` runs the model on synthetic VOC-shaped tensors (no dataset / checkpoint exists offline).

Batch size: the reference evaluates one image at a time (:212); nothing here depends on that, N images are
processed per call (BASELINE configs 2/3 use 8 and 32).
"""
import argparse
import types

import torch

from . import ops
from .ops import conv as conv_ops
from .net import resnet_v2, xception_body
from .net.variables import VariableStore
from .preprocessing import anchor_manipulator

# flag name -> default (light_head_rfcn_eval.py:41-139; data/summary/checkpoint flags kept for completeness)
_DEFAULTS = dict(
    num_readers=16, num_preprocessing_threads=48, num_cpu_threads=0, gpu_memory_fraction=1.,
    data_dir='../PASCAL/VOC_TF/VOC2007TEST_TF/', dataset_name='pascalvoc_2007', num_classes=21,
    model_dir='./logs_light/', log_every_n_steps=10, save_summary_steps=100, dataset_split_name='test',
    debug_dir='./Debug_light/', train_image_size=480, resnet_size=50, roi_one_image=64,
    data_format='channels_last',
    select_threshold=0.01, nms_threshold=0.3, nms_topk_percls=200, nms_topk=200, fg_ratio=0.25,
    match_threshold=0.53, neg_threshold_high=0.5, neg_threshold_low=0., rpn_anchors_per_image=256,
    rpn_pre_nms_top_n=5000, rpn_post_nms_top_n=1000, rpn_min_size=16 * 1. / 480, rpn_nms_thres=0.7,
    rpn_fg_ratio=0.5, rpn_match_threshold=0.7, rpn_neg_threshold=0.3, weight_decay=0.0002,
    checkpoint_path='./model/xception', model_scope='xception_lighthead', run_on_cloud=True,
    cloud_checkpoint_path='xception_model/xception_model.ckpt',
    # not a reference flag: which backbone builder to use ('xception' = the reference's XceptionBody,
    # 'resnet50' = the ResNet-50 v2 light-head composition of BASELINE configs 2/4)
    backbone='resnet50',
    # not a reference flag: 'f16x2' = the reference's precision (fp32 activations, split-operand tensor-core
    # convolutions with fp32-level error: within 1e-4 of the fp32 graph end to end; csrc/conv_gemm_f16x2.cu) -- what
    # bench.py reports; 'bf16' = the fast mode (~1e-2 per stage, 2.2x faster); 'fp32x3' = round 1's parity mode (tests)
    precision='bf16',
)
FLAGS = types.SimpleNamespace(**_DEFAULTS)
pool_method = 'max'  # light_head_rfcn_eval.py:172


def make_params(**overrides):
    p = dict(_DEFAULTS)
    p.update(overrides)
    return p


def input_pipeline(params, device="cuda"):
    """Anchor side of the reference's input_pipeline (:174-226): AnchorCreator over the stride-16 map,
    scales .2...8 + extra .1, ratios 1/2/.5, and the decode closures the model_fn receives through ``labels``."""
    size = params['train_image_size']
    fmap = size // 16
    creator = anchor_manipulator.AnchorCreator([size] * 2, layers_shapes=[(fmap, fmap)],
                                               anchor_scales=[[0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8]],
                                               extra_anchor_scales=[[0.1]], anchor_ratios=[[1., 2., .5]],
                                               layer_steps=[16])
    all_anchors, num_anchors_list = creator.get_all_anchors()
    enc = anchor_manipulator.AnchorEncoder(all_anchors, num_classes=params['num_classes'], allowed_borders=[0.],
                                           positive_threshold=params['rpn_match_threshold'],
                                           ignore_threshold=params['rpn_neg_threshold'],
                                           prior_scaling=[1., 1., 1., 1.], rpn_fg_thres=params['match_threshold'],
                                           rpn_bg_high_thres=params['neg_threshold_high'],
                                           rpn_bg_low_thres=params['neg_threshold_low'], device=device)
    return {'rpn_decode_fn': lambda pred: enc.decode_all_anchors([pred], squeeze_inner=True)[0],
            'head_decode_fn': lambda rois, pred: enc.ext_decode_rois(rois, pred, head_prior_scaling=[1., 1., 1., 1.]),
            'num_anchors_list': num_anchors_list, 'anchor_encoder': enc}


def det_min_size(image_shape, train_image_size, device):
    """utility/eval_helper.py:295 in fp32 on the host: max(0.0001, 0.03 * sqrt(float32(h*w) / (net_h*net_w))) per
    image -> fp32 tensor [N] on ``device`` (a host->device copy: build it outside CUDA-graph capture)."""
    import numpy as np
    shp = np.asarray(image_shape, dtype=np.int64).reshape(-1, 2)
    net = np.float32(int(train_image_size) * int(train_image_size))
    q = (shp[:, 0] * shp[:, 1]).astype(np.float32) / net
    ms = np.maximum(np.float32(0.0001), np.float32(0.03) * np.sqrt(q, dtype=np.float32)).astype(np.float32)
    return torch.from_numpy(ms).to(device)


def bboxes_eval(image_shape, bbox_img, cls_pred_prob, bboxes_pred, num_classes, params=None, min_size=None):
    """Detection post-processing of the reference's ``bboxes_eval`` (light_head_rfcn_eval.py:263-290, the
    '/device:CPU:0' block): per class c >= 1 select (score > select_threshold) -> clip to ``bbox_img`` -> filter
    (min size 0.03 of the net input, centre inside) -> resize to the original image frame -> top 2*nms_topk ->
    NMS(nms_threshold) -> nms_topk, zero padded.  On the GPU, batched over images AND classes (the reference handles
    one image per call on the CPU).

    image_shape [N,2] (h, w of the original images; host ints), bbox_img [N,4], cls_pred_prob [N,R,num_classes]
    (softmax scores, e.g. ``head_cls_score``), bboxes_pred [N,R,4] (decoded, e.g. ``bboxes_predict``).
    Returns ({c: scores [N,nms_topk]}, {c: bboxes [N,nms_topk,4]}) -- the reference's per-class dictionaries.
    TP/FP matching and mAP (``bboxes_matching_batch``, ``metrics``) are SURVEY 8 f3, not built."""
    p = params or _DEFAULTS
    dev = cls_pred_prob.device
    if min_size is None:  # ``min_size`` [N] fp32 on the device may be passed instead (det_min_size, precomputed)
        min_size = det_min_size(image_shape, p['train_image_size'], dev)
    scores, boxes = ops.det_postprocess(cls_pred_prob, bboxes_pred, bbox_img.to(dev, torch.float32), min_size,
                                        p['select_threshold'], 2 * p['nms_topk'], p['nms_topk'], p['nms_threshold'])
    return ({c: scores[:, c - 1] for c in range(1, num_classes)}, {c: boxes[:, c - 1] for c in range(1, num_classes)})


_SIDE_STREAMS = {}


def _side_stream(device):
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def lighr_head_model_fn(features, labels, mode, params, store=None, shuffle_keys=None):
    """features: images [N,3,H,W] fp32 CUDA (whitened, as light_head_preprocess_for_eval delivers them).
    Returns the ``predictions`` of the reference's model_fn plus the intermediate tensors tests compare."""
    assert mode in ("eval", "predict", "infer"), 'This script only support predict mode!'
    num_anchors = labels['num_anchors_list'][0]
    enc = labels['anchor_encoder']
    df = 'channels_last'
    main = torch.cuda.current_stream()
    side = _side_stream(features.device)
    st = {}

    def after_rpn_feat(rpn_feat_map):
        # RPN head on the main stream, then FORK: objectness/decode -> clip/filter/top-k -> NMS -> upsample run on
        # a second stream beside block_layer4 + large_sep_kernel (they only meet again at PsRoIAlign).  The
        # proposal kernels are one CTA per image, so the backbone convolutions leave that many SMs free meanwhile.
        st['rpn_out'] = xception_body.get_rpn(rpn_feat_map, num_anchors, False, df, 'rpn_head', store)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            # softmax[:, -1] + decode_all_anchors in one kernel (reference :389-399)
            st['score'], st['boxes'] = ops.rpn_decode(st['rpn_out'], 0, 2 * num_anchors, enc.device_anchors(0),
                                                      num_anchors)
            st['rois'], st['yxhw'], _ = ops.rpn_select(st['score'], st['boxes'], params['rpn_pre_nms_top_n'],
                                                       params['rpn_post_nms_top_n'], params['rpn_nms_thres'],
                                                       params['rpn_min_size'], shuffle_keys)
        conv_ops.MAX_CTAS = 148 - min(int(features.shape[0]), 16)

    with store.scope(params['model_scope']):
        try:
            if params.get('backbone', 'resnet50') == 'resnet50':
                rpn_feat_map, backbone_feat = resnet_v2.lighthead_resnet50_body(features, False, store,
                                                                                after_rpn_feat=after_rpn_feat)
            else:  # the reference's own backbone (light_head_rfcn_eval.py:383)
                rpn_feat_map, backbone_feat = xception_body.XceptionBody(features, params['num_classes'], False, df,
                                                                         store, after_mid=after_rpn_feat)
            large_sep_feature = xception_body.large_sep_kernel(backbone_feat, 256, 10 * 7 * 7, False, df,
                                                               'large_sep_feature', store)
        finally:
            conv_ops.MAX_CTAS = 0
        main.wait_stream(side)  # JOIN
        rpn_out, rpn_object_score, rpn_bboxes_pred = st['rpn_out'], st['score'], st['boxes']
        proposals_bboxes, yxhw = st['rois'], st['yxhw']
        for t in (rpn_object_score, rpn_bboxes_pred, proposals_bboxes, yxhw):
            t.record_stream(main)
        cls_score, bboxes_reg, head = xception_body.get_head(
            large_sep_feature, lambda input_, bboxes_, gw_, gh_: ops.ps_roi_align(input_, bboxes_, gw_, gh_, pool_method),
            7, 7, None, proposals_bboxes, params['num_classes'], False, False, 0, df, 'final_head', store,
            yxhw_bboxes=yxhw, return_fused=True)
        N, R = proposals_bboxes.shape[:2]
        head = head.reshape(N * R, -1)
        head_cls_score, head_bboxes_pred, classes, probabilities = ops.head_decode(
            proposals_bboxes.reshape(-1, 4), head, 0, params['num_classes'], params['num_classes'], with_classes=True)
    return {
        'classes': classes, 'probabilities': probabilities,
        'bboxes_predict': head_bboxes_pred, 'head_cls_score': head_cls_score,
        # intermediates (not in the reference's predictions dict; used by the parity tests)
        'rpn_feat_map': rpn_feat_map, 'backbone_feat': backbone_feat, 'rpn_out': rpn_out,
        'large_sep_feature': large_sep_feature, 'rpn_object_score': rpn_object_score,
        'rpn_bboxes_pred': rpn_bboxes_pred, 'proposals_bboxes': proposals_bboxes, 'cls_score': cls_score,
        'bboxes_reg': bboxes_reg,
    }


class LightHeadRFCN(object):
    """Convenience wrapper: variables + anchors + one call per batch (what ``tf.estimator.Estimator.predict``
    drives in the reference, :447-503)."""

    def __init__(self, params=None, seed=0, device="cuda", state_dict=None):
        self.params = params or make_params()
        self.store = VariableStore(device=device, seed=seed, state_dict=state_dict)
        self.labels = input_pipeline(self.params, device=device)
        self._det_consts = {}
        self._check_restored = False  # from_checkpoint: report variables the checkpoint did not provide

    @classmethod
    def from_checkpoint(cls, params=None, checkpoint_path=None, checkpoint_model_scope=None, seed=0, device="cuda"):
        """Build the model from a TensorFlow V2 checkpoint of the reference (``--checkpoint_path`` / ``--model_dir``,
        light_head_rfcn_eval.py:125-137, :499): variables found in the checkpoint are used, the others keep their
        seeded initial values (utility/train_helper.py)."""
        from .utility import train_helper
        params = params or make_params()
        sd = train_helper.checkpoint_to_state_dict(checkpoint_path or params['checkpoint_path'], params['model_scope'],
                                                   checkpoint_model_scope)
        model = cls(params, seed=seed, device=device, state_dict=sd)
        model._check_restored = True
        return model

    def det_consts(self, n, image_shape=None, bbox_img=None, device="cuda"):
        """(bbox_img [n,4], min_size [n]) device tensors for the post-processing; cached, so a CUDA-graph capture of
        ``__call__(..., detections=True)`` finds them already on the device.  Defaults: the whole net input is the
        image (bbox_img = [0,0,1,1], image_shape = train_image_size squared) -- the synthetic-tensor case."""
        size = self.params['train_image_size']
        shp = tuple(map(tuple, image_shape)) if image_shape is not None else ((size, size),) * n
        ref = tuple(map(tuple, bbox_img)) if bbox_img is not None else ((0., 0., 1., 1.),) * n
        key = (n, shp, ref, str(device))
        if key not in self._det_consts:
            self._det_consts[key] = (torch.tensor(ref, dtype=torch.float32, device=device),
                                     det_min_size(shp, size, device))
        return self._det_consts[key]

    def __call__(self, images, shuffle_keys=None, detections=False, image_shape=None, bbox_img=None):
        """One batch through the model_fn.  ``detections=True`` also runs ``bboxes_eval`` and adds 'det_scores'
        [N,num_classes-1,nms_topk] and 'det_bboxes' [N,num_classes-1,nms_topk,4] (class c at index c-1)."""
        # a fresh naming pass per call: variables are looked up by the same automatic names every time
        self.store._counters = [{}]
        with conv_ops.precision(self.params.get('precision', 'bf16')):
            out = lighr_head_model_fn(images, self.labels, "eval", self.params, store=self.store,
                                      shuffle_keys=shuffle_keys)
        if self._check_restored:
            self._check_restored = False
            missing = list(self.store.created)
            if missing:
                import warnings
                stats = [k for k in missing if k.endswith(("moving_mean", "moving_variance", "gamma", "beta"))]
                warnings.warn("from_checkpoint: %d variable(s) of the model were not in the checkpoint and keep their "
                              "initial values (%d of them batch-norm scale / shift / moving statistics, initialised to "
                              "identity), e.g. %s" % (len(missing), len(stats), ", ".join(missing[:4])), RuntimeWarning)
        if detections:
            N, C = images.shape[0], self.params['num_classes']
            ref, min_size = self.det_consts(N, image_shape, bbox_img, images.device)
            p = self.params
            out['det_scores'], out['det_bboxes'] = ops.det_postprocess(
                out['head_cls_score'].reshape(N, -1, C), out['bboxes_predict'].reshape(N, -1, 4), ref, min_size,
                p['select_threshold'], 2 * p['nms_topk'], p['nms_topk'], p['nms_threshold'])
        return out


def arg_parser():
    """The reference's flags (light_head_rfcn_eval.py:41-139), same names and defaults, as ``--name value``."""
    ap = argparse.ArgumentParser(description="Light-Head R-CNN inference on synthetic VOC-shaped tensors")
    for k, v in _DEFAULTS.items():
        if isinstance(v, bool):
            ap.add_argument("--" + k, type=lambda s: s.lower() in ("1", "true", "yes"), default=v)
        else:
            ap.add_argument("--" + k, type=type(v), default=v)
    ap.add_argument("--batch_size", type=int, default=8)
    return ap


def main(argv=None):
    args = arg_parser().parse_args(argv)
    params = make_params(**{k: getattr(args, k) for k in _DEFAULTS})
    model = LightHeadRFCN(params)
    g = torch.Generator(device="cuda").manual_seed(1)
    size = params['train_image_size']
    images = torch.rand((args.batch_size, 3, size, size), generator=g, device="cuda") * 2 - 1
    out = model(images)
    torch.cuda.synchronize()
    print("classes", out['classes'][:10].tolist())
    print("probabilities", [round(float(p), 4) for p in out['probabilities'][:10]])
    print("bboxes_predict", out['bboxes_predict'][:3].tolist())


if __name__ == '__main__':
    main()
