"""``ps_roi_align`` / ``ps_roi_align_grad`` -- the reference's custom-op module surface.

The reference loads ``libps_roi_align.so`` with ``tf.load_op_library`` and calls
``op_module.ps_roi_align(inputs, rois, grid_dim_width, grid_dim_height, pool_method)
-> (pooled_features, pooled_index)`` and ``op_module.ps_roi_align_grad(inputs, rois,
pooled_features_grad, pooled_index, grid_dim_width, grid_dim_height, pool_method) -> grad_output``
(light_head_rfcn_train.py:187-213; op defs cpp/PSROIPooling/ps_roi_align_op.cc:38-48,
ps_roi_align_grad_op.cc:39-48).  Same names, argument order, shapes and dtypes here; the
validation errors of ``PSROIAlignOp`` (ps_roi_align_op.cc:208-226) surface as ``ValueError``.

Tensors are CUDA torch tensors (device memory + stream only); the work is done by
``xdet_psroi_align_fwd`` / ``xdet_psroi_align_bwd`` in ``csrc/psroi_align.cu``.
"""
import torch

from .. import _native

VARIANTS = {"auto": 0, "gather": 1, "planes": 2, "select": 3}


def _use_max(pool_method):
    # ps_roi_align_op.cc:216 -- the attr must contain "mean" or "max"; "max" wins if both (:242)
    if not isinstance(pool_method, str) or ("mean" not in pool_method and "max" not in pool_method):
        raise ValueError("Need Attr pool_method to be either 'mean' or 'max', got %r" % (pool_method,))
    return 1 if "max" in pool_method else 0


def _check_inputs(inputs, rois, gw, gh):
    if gw < 0 or gh < 0:
        raise ValueError("Need Attr grid_dim_width/grid_dim_height >= 0, got %d, %d" % (gw, gh))
    if inputs.dim() != 4:
        raise ValueError("inputs must be in 'NCHW' format.")
    if rois.dim() != 3 or rois.shape[2] != 4:
        raise ValueError("rois must be in 'batch_size x num_rois x 4' format.")
    if inputs.shape[0] != rois.shape[0]:
        raise ValueError("'batch_size' in inputs and rois don't match.")
    if not (inputs.is_cuda and rois.is_cuda):
        raise ValueError("ps_roi_align runs on the GPU only (no CPU fallback): pass CUDA tensors")
    if inputs.dtype != torch.float32 or rois.dtype != torch.float32:
        raise ValueError("ps_roi_align supports T in {float} only (ps_roi_align_op.cc:39)")


def _stream():
    return torch.cuda.current_stream().cuda_stream


def ps_roi_align(inputs, rois, grid_dim_width, grid_dim_height, pool_method, variant="auto"):
    """inputs [N,C,H,W] f32, rois [N,R,4] f32 (cy,cx,h,w in [0,1]) ->
    (pooled_features [N,R,gw*gh,C/(gw*gh)] f32, pooled_index same shape i32)."""
    gw, gh = int(grid_dim_width), int(grid_dim_height)
    use_max = _use_max(pool_method)
    _check_inputs(inputs, rois, gw, gh)
    inputs = inputs.contiguous()
    rois = rois.contiguous()
    N, C, H, W = inputs.shape
    R = rois.shape[1]
    G = gw * gh
    if G == 0 or C % G != 0:
        raise ValueError("channels (%d) must be divisible by grid_dim_width*grid_dim_height (%d)" % (C, G))
    with torch.cuda.device(inputs.device):
        pooled = torch.empty((N, R, G, C // G), dtype=torch.float32, device=inputs.device)
        index = torch.empty((N, R, G, C // G), dtype=torch.int32, device=inputs.device)
        rc = _native.lib().xdet_psroi_align_fwd_ex(inputs.data_ptr(), rois.data_ptr(), pooled.data_ptr(),
                                                   index.data_ptr(), N, C, H, W, R, gw, gh, use_max,
                                                   VARIANTS[variant], _stream())
    _native.check(rc)
    return pooled, index


def ps_roi_align_grad(inputs, rois, pooled_features_grad, pooled_index, grid_dim_width, grid_dim_height,
                      pool_method):
    """-> grad_output with the shape of ``inputs`` (only the shape of ``inputs`` is used, as in
    the reference: ps_roi_align_grad_op.cc:329-353)."""
    gw, gh = int(grid_dim_width), int(grid_dim_height)
    use_max = _use_max(pool_method)
    _check_inputs(inputs, rois, gw, gh)
    N, C, H, W = inputs.shape
    R = rois.shape[1]
    G = gw * gh
    if G == 0 or C % G != 0:
        raise ValueError("channels (%d) must be divisible by grid_dim_width*grid_dim_height (%d)" % (C, G))
    if tuple(pooled_features_grad.shape) != tuple(pooled_index.shape):
        raise ValueError("pooled_index and pooled_features_grad must have the same shape")
    if tuple(pooled_features_grad.shape) != (N, R, G, C // G):
        raise ValueError("both pooled_index and pooled_features_grad must have the shape "
                         "'batch_size x num_rois x grid_size x bank_size'")
    rois = rois.contiguous()
    g = pooled_features_grad.contiguous().float()
    idx = pooled_index.contiguous().to(torch.int32)
    with torch.cuda.device(inputs.device):
        grad = torch.empty((N, C, H, W), dtype=torch.float32, device=inputs.device)
        rc = _native.lib().xdet_psroi_align_bwd(rois.data_ptr(), g.data_ptr(), idx.data_ptr(), grad.data_ptr(),
                                                N, C, H, W, R, gw, gh, use_max, _stream())
    _native.check(rc)
    return grad


class PsRoiAlign(torch.autograd.Function):
    """Autograd pairing of the two ops -- what ``@ops.RegisterGradient("PsRoiAlign")`` does at
    light_head_rfcn_train.py:201-213 (gradient flows to ``inputs`` only)."""

    @staticmethod
    def forward(ctx, inputs, rois, grid_dim_width, grid_dim_height, pool_method):
        pooled, index = ps_roi_align(inputs, rois, grid_dim_width, grid_dim_height, pool_method)
        ctx.save_for_backward(inputs, rois, index)
        ctx.attrs = (grid_dim_width, grid_dim_height, pool_method)
        ctx.mark_non_differentiable(index)
        return pooled, index

    @staticmethod
    def backward(ctx, grad_pooled, _grad_index):
        inputs, rois, index = ctx.saved_tensors
        gw, gh, method = ctx.attrs
        return ps_roi_align_grad(inputs, rois, grad_pooled, index, gw, gh, method), None, None, None, None
