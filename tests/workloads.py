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


def adversarial_maps():
    """Planes built to defeat an approximate arg-max: exact ties, near-ties, zeros, huge and tiny magnitudes,
    non-finite values.  SELECT must fall back to the exact loop wherever the fp32 pass cannot prove the winner."""
    rng = np.random.default_rng(99)
    base = rng.standard_normal((1, 98, 30, 30), dtype=np.float32)
    relu = np.maximum(base, 0)                                   # ~50 % exact zeros (the model's thin map is post-ReLU)
    sparse = np.where(rng.random(base.shape) < 0.9, 0, base).astype(np.float32)
    const = np.full_like(base, 0.7)                              # every sample ties
    negzero = np.where(rng.random(base.shape) < 0.5, -0.0, 0.0).astype(np.float32)
    quant = np.round(base * 2).astype(np.float32) / 2            # few distinct values: many exact ties
    near = (1.0 + rng.integers(0, 4, base.shape) * np.float32(2.0 ** -23)).astype(np.float32)  # 1-ulp steps
    huge = (base * np.float32(1e30)).astype(np.float32)
    tiny = (base * np.float32(1e-38)).astype(np.float32)         # subnormal products
    mixed = base.copy()
    mixed[:, ::7] *= np.float32(1e20)
    nonfin = base.copy()
    nonfin[0, 3, 4, 5] = np.nan
    nonfin[0, 10, 20, 11] = np.inf
    nonfin[0, 11, 2, 7] = -np.inf
    ramp = np.broadcast_to(np.arange(30, dtype=np.float32)[None, None, None, :], base.shape).copy()  # affine in x
    return {"relu": relu, "sparse": sparse, "const": const, "negzero": negzero, "quant": quant, "near": near,
            "huge": huge, "tiny": tiny, "mixed": mixed, "nonfinite": nonfin, "ramp": ramp}
