from .ps_roi_align import ps_roi_align, ps_roi_align_grad, PsRoiAlign  # noqa: F401
