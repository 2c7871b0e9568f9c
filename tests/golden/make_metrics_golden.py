#!/usr/bin/env python
"""Mint goldens for the streaming detection metrics from the reference's own utility/metrics.py, run unmodified under
the numpy TensorFlow stand-in, chained as light_head_rfcn_eval.bboxes_eval chains them (:291-323):
streaming_tp_fp_arrays (dict form, one call per image = one run of its update ops), then per class precision_recall,
average_precision_voc07 / voc12, and the mAPs.  Run in the build container only.
    python tests/golden/make_metrics_golden.py"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "oracle", "tf_shim"), "/root/reference", ROOT]

import numpy as np  # noqa: E402
import tensorflow as tf  # noqa: E402  (the stand-in)

from utility import metrics  # noqa: E402  (reference)

CLASSES, IMAGES, K = (1, 2, 3, 4), 7, 12


def main():
    assert "tf_shim" in tf.__file__
    rs = np.random.RandomState(99)
    out = {"classes": np.array(CLASSES), "images": np.int64(IMAGES)}
    tp_fp_metric = None
    for i in range(IMAGES):
        num_gbboxes, tp, fp, scores = {}, {}, {}, {}
        for c in CLASSES:
            s = np.sort(rs.uniform(0, 1, K).astype(np.float32))[::-1].copy()
            s[rs.randint(6, K):] = 0            # the zero padding of bboxes_nms_batch
            s[rs.randint(0, K)] = 5e-5          # below the 1e-4 removal threshold
            kind = rs.randint(0, 3, K)          # 0: TP, 1: FP, 2: neither (matched a 'difficult' object)
            if c == 4:
                kind[:] = 1                     # a class that is never detected correctly
            t, f = kind == 0, kind == 1
            n = np.int64(t.sum() + rs.randint(0, 3)) if c != 4 else np.int64(rs.randint(0, 2))
            num_gbboxes[c], tp[c], fp[c], scores[c] = tf.constant(n), tf.constant(t), tf.constant(f), tf.constant(s)
            out["in_%d_%d_n" % (i, c)], out["in_%d_%d_tp" % (i, c)] = n, t
            out["in_%d_%d_fp" % (i, c)], out["in_%d_%d_scores" % (i, c)] = f, s
        tp_fp_metric = metrics.streaming_tp_fp_arrays(num_gbboxes, tp, fp, scores)
    aps_voc07, aps_voc12 = {}, {}
    for c in tp_fp_metric[0].keys():
        vals = tp_fp_metric[0][c]
        for name, v in zip(('nobjects', 'ndetections', 'tp', 'fp', 'scores'), vals):
            out["acc_%d_%s" % (c, name)] = np.asarray(v)
        with np.errstate(all="ignore"):
            prec, rec = metrics.precision_recall(*vals)
            out["prec_%d" % c], out["rec_%d" % c] = np.asarray(prec), np.asarray(rec)
            aps_voc07[c] = metrics.average_precision_voc07(prec, rec)
            aps_voc12[c] = metrics.average_precision_voc12(prec, rec)
        out["ap07_%d" % c], out["ap12_%d" % c] = np.float64(aps_voc07[c]), np.float64(aps_voc12[c])
    out["map07"] = np.float64(tf.add_n(list(aps_voc07.values())) / len(aps_voc07))
    out["map12"] = np.float64(tf.add_n(list(aps_voc12.values())) / len(aps_voc12))
    print({k: float(out[k]) for k in out if k.startswith(("ap", "map"))})
    print({c: (int(out["acc_%d_nobjects" % c]), int(out["acc_%d_ndetections" % c])) for c in CLASSES})
    path = os.path.join(HERE, "metrics_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
