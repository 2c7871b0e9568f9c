"""CPU: the division-free fast path of the NMS IoU test (csrc/rpn_proposals.cu `iou_greater`) decides exactly like the
reference's `inter / union > thr` (one fp32 division, TF r1.6 NonMaxSuppressionV2).  The kernel skips the division
when inter lies outside [t*(1-1e-6), t*(1+1e-6)], t = fl(thr*union); this test replays both forms in numpy fp32 on
random pairs and on pairs constructed to sit within a few ulps of the threshold."""
import numpy as np

F = np.float32


def decide_reference(inter, uni, thr):
    with np.errstate(divide="ignore", invalid="ignore"):
        return (inter / uni).astype(F) > F(thr)


def decide_kernel(inter, uni, thr):
    t = (F(thr) * uni).astype(F)
    hi = (t * F(1.000001)).astype(F)
    lo = (t * F(0.999999)).astype(F)
    out = decide_reference(inter, uni, thr)          # the slow path: the same division
    fast = (uni > 0) & (F(thr) >= 0)
    out = np.where(fast & (inter <= 0), False, out)
    out = np.where(fast & (inter > 0) & (inter > hi), True, out)
    out = np.where(fast & (inter > 0) & ~(inter > hi) & (inter < lo), False, out)
    return out


def test_bracket_agrees_with_division():
    rng = np.random.default_rng(0)
    for thr in (0.7, 0.3, 0.5, 0.0, 1.0):
        uni = rng.uniform(1e-6, 2.0, 2_000_000).astype(F)
        inter = (uni * rng.uniform(0, 1, uni.shape).astype(F)).astype(F)
        assert np.array_equal(decide_kernel(inter, uni, thr), decide_reference(inter, uni, thr))
        # adversarial: inter within +-64 ulps of thr*union
        base = (F(thr) * uni).astype(F)
        for k in range(-64, 65, 3):
            near = (base.view(np.int32) + np.int32(k)).view(F)
            near = np.where(np.isfinite(near) & (near >= 0), near, base)
            assert np.array_equal(decide_kernel(near, uni, thr), decide_reference(near, uni, thr)), (thr, k)
