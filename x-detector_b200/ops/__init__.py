from .ps_roi_align import ps_roi_align, ps_roi_align_grad, PsRoiAlign  # noqa: F401
from .conv import conv2d_nhwc, linear, pack_conv_weight, same_pad  # noqa: F401
