from .ps_roi_align import ps_roi_align, ps_roi_align_grad, PsRoiAlign  # noqa: F401
from .conv import conv2d_dgrad, conv2d_nhwc, conv2d_image_fold, conv2d_wgrad, pack_dgrad_weight, linear, pack_conv_weight, pack_fold_weight, same_pad, split3, f32_post, precision, split3_values, split2, pair_of, PairWeight  # noqa: F401
from .layout import affine_relu, depthwise3x3, f32_to_bf16_rows, im2col, image_to_nhwc8, maxpool3x3s2_same  # noqa: F401
from .proposals import det_postprocess, head_decode, rpn_decode, rpn_select  # noqa: F401
from . import train  # noqa: F401,E402
