"""Mint golden vectors by RUNNING THE REFERENCE ITSELF (its unmodified C++ CPU op, compiled by
oracle/Makefile against oracle/ref_shim) -- run in the build container where /root/reference is
mounted:   python tests/golden/make_golden.py
Outputs (committed): tests/golden/psroi_golden.npz
  * the reference's only fixture (cpp/PSROIPooling/test_op.py:52-81): 1x16x5x5, 3 RoIs, 2x2, mean+max,
    forward outputs and the gradient of an all-ones upstream;
  * a small seeded random case with full outputs (fwd + bwd, mean + max);
  * SHA-256 digests of the reference outputs for the config-1 "S-model" case
    (1x490x30x30, 300+4 RoIs, 7x7), so the GPU path can be pinned to the reference
    on the GPU box, where /root/reference does not exist.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import psroi  # noqa: E402
from tests import workloads  # noqa: E402


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    psroi.build(ref=True)
    out = {}
    # 1. the reference's own fixture
    x = np.tile(np.arange(1, 26, dtype=np.float32).reshape(1, 1, 5, 5), (1, 16, 1, 1)).copy()
    rois = np.array([[[0.2, 0.2, 0.7, 0.7], [0.5, 0.5, 0.9, 0.9], [0.9, 0.9, 1.0, 1.0]]], dtype=np.float32)
    out["fix_inputs"], out["fix_rois"] = x, rois
    for m in ("mean", "max"):
        p, i = psroi.psroi_align_fwd(x, rois, 2, 2, m, impl="ref")
        g = psroi.psroi_align_bwd(x.shape, rois, np.ones_like(p), i, 2, 2, m, impl="ref")
        out["fix_%s_pooled" % m], out["fix_%s_index" % m], out["fix_%s_grad" % m] = p, i, g
    # 2. small random case, full outputs
    x = workloads.make_map(2, 98, 30, 30, seed=10)
    rois = workloads.make_rois(2, 28, seed=11, edge_cases=True)
    out["small_seed"] = np.array([10, 11])
    for m in ("mean", "max"):
        p, i = psroi.psroi_align_fwd(x, rois, 7, 7, m, impl="ref")
        i[workloads.degenerate_mask(rois)] = 0
        gup = np.random.default_rng(12).standard_normal(p.shape, dtype=np.float32)
        g = psroi.psroi_align_bwd(x.shape, rois, gup, i, 7, 7, m, impl="ref")
        out["small_%s_pooled" % m], out["small_%s_index" % m], out["small_%s_grad" % m] = p, i, g
    # 3. config-1 S-model digests
    x = workloads.make_map(1, 490, 30, 30, seed=0)
    rois = workloads.make_rois(1, 300, seed=0, edge_cases=True)
    for m in ("mean", "max"):
        p, i = psroi.psroi_align_fwd(x, rois, 7, 7, m, impl="ref")
        i[workloads.degenerate_mask(rois)] = 0
        gup = np.random.default_rng(1).standard_normal(p.shape, dtype=np.float32)
        g = psroi.psroi_align_bwd(x.shape, rois, gup, i, 7, 7, m, impl="ref")
        out["c1_%s_sha" % m] = np.array([digest(p), digest(i), digest(g)])
    np.savez_compressed(os.path.join(HERE, "psroi_golden.npz"), **out)
    print("wrote psroi_golden.npz:", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
