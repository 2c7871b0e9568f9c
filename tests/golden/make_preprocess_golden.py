#!/usr/bin/env python
"""Mint goldens for the eval input pipeline from the reference's own preprocessing/common_preprocessing.py
(light_head_preprocess_for_eval :383-440 and light_head_preprocess_for_test :443-458, WARP_RESIZE, NCHW), run
unmodified under the numpy TensorFlow stand-in.  The stand-in's tf.image.resize_images is TF r1.6's ResizeBilinear
(align_corners=False) written as two float64 interpolation matrices -- a different formulation from the fp32
lerp-by-lerp restatement in oracle/preprocess.py and in the kernel, so the sampling grid, the edge clamp and the
whitening constants are cross-checked rather than repeated.  Run in the build container only.
    python tests/golden/make_preprocess_golden.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), "/root/reference", ROOT]

import numpy as np  # noqa: E402
import tensorflow as tf  # noqa: E402  (the stand-in)

from preprocessing import common_preprocessing as cp  # noqa: E402  (reference)

CASES = [((37, 53), (48, 64)), ((61, 45), (32, 24)), ((40, 40), (40, 40)), ((23, 90), (64, 64))]


def main():
    assert "tf_shim" in tf.__file__
    out = {}
    rs = np.random.RandomState(7)
    for i, (shape, size) in enumerate(CASES):
        img = rs.randint(0, 256, shape + (3,)).astype(np.uint8)
        out["img_%d" % i], out["size_%d" % i] = img, np.array(size)
        out["test_%d" % i] = np.asarray(cp.light_head_preprocess_for_test(tf.constant(img), list(size), data_format='NCHW'))
    labels = np.array([3, 7, 9, 12], np.int64)
    bboxes = rs.uniform(0, 1, (4, 4)).astype(np.float32)
    difficults = np.array([0, 1, 0, 1], np.int64)
    image, lab, bb, bbox_img = cp.light_head_preprocess_for_eval(tf.constant(out["img_0"]), tf.constant(labels),
                                                                 tf.constant(bboxes), out_shape=list(CASES[0][1]),
                                                                 data_format='NCHW', difficults=tf.constant(difficults))
    out.update(eval_labels_in=labels, eval_bboxes_in=bboxes, eval_difficults=difficults, eval_image=np.asarray(image),
               eval_labels=np.asarray(lab), eval_bboxes=np.asarray(bb), eval_bbox_img=np.asarray(bbox_img))
    path = os.path.join(HERE, "preprocess_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
