"""Eval / test input pipeline -- the API of the reference's ``preprocessing/common_preprocessing.py``
(``light_head_preprocess_for_eval`` :383-441, ``light_head_preprocess_for_test`` :443-458) for the scripts' default
``Resize.WARP_RESIZE``: whitening + bilinear resize + NCHW transpose as ONE GPU kernel per image
(``xdet_preprocess_eval_u8``).  JPEG decoding and TFRecord parsing stay upstream (out of scope).
"""
import ctypes
from enum import IntEnum

import torch

from .. import _native

_R_MEAN = 123.68  # preprocessing/common_preprocessing.py:36-38
_G_MEAN = 116.78
_B_MEAN = 103.94
EVAL_SIZE = (320, 320)  # :44
# :30-33; only the warp resize (what both scripts use) has a kernel
Resize = IntEnum('Resize', ('NONE', 'CENTRAL_CROP', 'PAD_AND_RESIZE', 'WARP_RESIZE'))


def _means():
    return (ctypes.c_float * 3)(_R_MEAN / 127.5, _G_MEAN / 127.5, _B_MEAN / 127.5)


def light_head_preprocess_for_test(image, out_shape, data_format='NCHW', resize=Resize.WARP_RESIZE,
                                   scope='light_head_preprocessing_test', out=None):
    """image: uint8 [H,W,3] RGB CUDA tensor -> fp32 [3,out_h,out_w] (``data_format='NCHW'``, what the model_fn takes;
    'NHWC' returns the transposed view).  ``out``: optional [3,out_h,out_w] slice of a batch tensor to write into."""
    if image.dim() != 3 or image.shape[2] != 3:
        raise ValueError('Input must be of size [height, width, C>0]')
    if resize != Resize.WARP_RESIZE:
        raise NotImplementedError('only Resize.WARP_RESIZE (the mode the train / eval scripts use) is built')
    assert image.dtype == torch.uint8 and image.is_cuda
    image = image.contiguous()
    H, W, _ = image.shape
    Ho, Wo = int(out_shape[0]), int(out_shape[1])
    if out is None:
        out = torch.empty((3, Ho, Wo), dtype=torch.float32, device=image.device)
    assert out.shape == (3, Ho, Wo) and out.dtype == torch.float32 and out.is_contiguous()
    rc = _native.lib().xdet_preprocess_eval_u8(image.data_ptr(), H, W, Ho, Wo, ctypes.byref(_means()), out.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream)
    _native.check(rc)
    return out if data_format == 'NCHW' else out.permute(1, 2, 0)


def light_head_preprocess_for_eval(image, labels, bboxes, out_shape=EVAL_SIZE, data_format='NCHW', difficults=None,
                                   resize=Resize.WARP_RESIZE, scope='light_head_preprocessing_eval', out=None):
    """Reference signature (:383-441, resize = WARP_RESIZE).  Returns (image, labels, bboxes, bbox_img): with a warp
    resize the whole net input is the image, so ``bbox_img`` = [0, 0, 1, 1] and the (normalised) boxes are unchanged;
    'difficult' objects are removed from labels / bboxes when ``difficults`` is given (:433-436)."""
    img = light_head_preprocess_for_test(image, out_shape, data_format, resize=resize, out=out)
    bbox_img = torch.tensor([0., 0., 1., 1.], dtype=torch.float32, device=image.device)
    if difficults is not None and labels is not None:
        mask = ~difficults.to(torch.bool)
        labels, bboxes = labels[mask], bboxes[mask]
    return img, labels, bboxes, bbox_img
