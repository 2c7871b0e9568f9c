"""CPU oracle (TEST INFRASTRUCTURE ONLY) of the eval input pipeline: numpy fp32 restatement of
``light_head_preprocess_for_test`` (preprocessing/common_preprocessing.py:443-458, WARP_RESIZE) with TF r1.6's
ResizeBilinear (legacy sampling, align_corners=False; tensorflow/core/kernels/resize_bilinear_op.cc, restated from
its published algorithm: TF is not installable offline).  Pinned around that kernel by the reference's own
preprocessing/common_preprocessing.py run under the numpy TensorFlow stand-in (tests/golden/make_preprocess_golden.py,
tests/test_preprocess_golden.py): dtype conversion, whitening constants, warp resize, NCHW transpose, difficult-box
removal, bbox_img; the stand-in's resize is an independent float64 matrix formulation of the same sampling rule."""
import numpy as np

F = np.float32
MEANS = np.array([123.68 / 127.5, 116.78 / 127.5, 103.94 / 127.5], F)


def preprocess_for_test(image_u8, out_shape):
    """image_u8 [H,W,3] uint8 -> [3,Ho,Wo] float32."""
    H, W, _ = image_u8.shape
    Ho, Wo = out_shape
    x = (image_u8.astype(F) * F(1.0 / 255.0)).astype(F)
    x = (x * F(2.0)).astype(F)
    x = (x - MEANS).astype(F)
    sy, sx = F(F(H) / F(Ho)), F(F(W) / F(Wo))
    in_y = (np.arange(Ho, dtype=F) * sy).astype(F)
    in_x = (np.arange(Wo, dtype=F) * sx).astype(F)
    y0 = in_y.astype(np.int64)
    x0 = in_x.astype(np.int64)
    y1 = np.minimum(y0 + 1, H - 1)
    x1 = np.minimum(x0 + 1, W - 1)
    yl = (in_y - y0.astype(F)).astype(F)[:, None, None]
    xl = (in_x - x0.astype(F)).astype(F)[None, :, None]
    tl, tr = x[y0][:, x0], x[y0][:, x1]
    bl, br = x[y1][:, x0], x[y1][:, x1]
    top = (tl + ((tr - tl).astype(F) * xl).astype(F)).astype(F)
    bot = (bl + ((br - bl).astype(F) * xl).astype(F)).astype(F)
    out = (top + ((bot - top).astype(F) * yl).astype(F)).astype(F)
    return np.ascontiguousarray(out.transpose(2, 0, 1))
