"""Seeded synthetic inputs shared by tests, bench.py and the golden generator."""
import numpy as np


def make_rois(n, r, seed, min_side=16.0 / 480.0, max_side=0.9, edge_cases=False):
    """RoIs as the detector produces them: boxes clipped to [0,1] (net/xception_body.py:173-194)
    with both sides > rpn_min_size, converted to (cy, cx, h, w) with the fp32 op order of
    `_point2center` (net/xception_body.py:215-218)."""
    rng = np.random.default_rng(seed)
    cy = rng.uniform(0.05, 0.95, (n, r))
    cx = rng.uniform(0.05, 0.95, (n, r))
    h = rng.uniform(min_side, max_side, (n, r))
    w = rng.uniform(min_side, max_side, (n, r))
    f = np.float32
    ymin = np.clip(cy - h / 2, 0, 1).astype(f)
    ymax = np.clip(cy + h / 2, 0, 1).astype(f)
    xmin = np.clip(cx - w / 2, 0, 1).astype(f)
    xmax = np.clip(cx + w / 2, 0, 1).astype(f)
    hh = ymax - ymin
    ww = xmax - xmin
    rois = np.stack([ymin + hh / f(2), xmin + ww / f(2), hh, ww], -1).astype(f)
    if edge_cases:
        # full image, ~1 pixel, touching the far border, degenerate (h = 0)
        extra = np.array([[0.5, 0.5, 1.0, 1.0], [0.4, 0.6, 1.0 / 30, 1.0 / 30], [0.9, 0.95, 0.2, 0.1],
                          [0.5, 0.5, 0.0, 0.3]], dtype=f)
        rois = np.concatenate([rois, np.broadcast_to(extra, (n,) + extra.shape)], axis=1)
    return np.ascontiguousarray(rois)


def make_map(n, c, h, w, seed):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, c, h, w), dtype=np.float32)


def degenerate_mask(rois):
    """True where the reference leaves pooled_index unwritten (ps_roi_align_op.cc:123-126)."""
    tiny = np.finfo(np.float32).tiny
    return (rois[..., 2] < tiny) | (rois[..., 3] < tiny)
