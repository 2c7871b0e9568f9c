"""CPU: the error bound behind the PsRoIAlign SELECT kernel (csrc/psroi_align.cu), checked numerically.

The kernel proves the reference's arg-max sample from fp32 "keys": K_s = the separable fp32 blend of sample s with the
sample id written over its 6 low mantissa bits.  Its decision rule relies on
        |K_s - V_s| <= alpha * cm + beta * |K_s|,   alpha = 8 * 2^-24, beta = 63 * 2^-23, cm = max|plane|,
where V_s is the reference's value (fp64 blend rounded to fp32, ps_roi_align_op.cc:173-176).  This test restates both
computations in numpy (same operations, same order) on random and adversarial planes and asserts the bound -- and that
the rule "gap > T  =>  unique exact maximum" never fires wrongly."""
import numpy as np

F = np.float32
D = np.float64
ALPHA = D(8 * 2.0 ** -24)
BETA = D(63 * 2.0 ** -23)


def fma32(a, b, c):
    """fp32 fused multiply-add, emulated through float64 (exact product; the sum may double-round in rare cases, which
    only loosens the emulation by < 1 ulp -- far inside the margins tested)."""
    return (a.astype(D) * b.astype(D) + c.astype(D)).astype(F)


def roi_samples(roi, H, W, gw, gh, row, col):
    """Sample coordinates of one bin, fp32 op order of ps_roi_align_op.cc:129-164 (in-contract RoIs)."""
    r0, r1, r2, r3 = (F(v) for v in roi)
    cy, cx = F(r0 * F(H)), F(r1 * F(W))
    rh, rw = max(F(r2 * F(H)), F(1)), max(F(r3 * F(W)), F(1))
    hh, hw = F(rh * F(0.5)), F(rw * F(0.5))
    ymin, xmin = max(F(cy - hh), F(0)), max(F(cx - hw), F(0))
    ymax, xmax = min(F(cy + hh), F(H)), min(F(cx + hw), F(W))
    bin_w, bin_h = F(F(xmax - xmin) / F(gw)), F(F(ymax - ymin) / F(gh))
    nw, nh = int(bin_w) + 1, int(bin_h) + 1
    sw, sh = F(bin_w / F(nw)), F(bin_h / F(nh))
    x0, y0 = F(xmin + F(bin_w * F(col))), F(ymin + F(bin_h * F(row)))
    xs = np.array([F(F(x0 + F(sw * F(i))) + F(sw * F(0.5))) for i in range(nw)], F)
    ys = np.array([F(F(y0 + F(sh * F(i))) + F(sh * F(0.5))) for i in range(nh)], F)
    return ys, xs


def keys_and_values(plane, ys, xs):
    """-> (K [nh*nw] fp32 keys as the kernel forms them, V [nh*nw] fp32 reference values), h-major sample order."""
    H, W = plane.shape
    K, V = [], []
    for hi, y in enumerate(ys):
        iy = int(y)
        fy = F(y - F(iy))
        iy0, iy1 = min(iy, H - 1), min(iy + 1, H - 1)
        ay = F(F(1) - fy)
        for wi, x in enumerate(xs):
            ix = int(x)
            fx = F(x - F(ix))
            ix0, ix1 = min(ix, W - 1), min(ix + 1, W - 1)
            p00, p01, p10, p11 = plane[iy0, ix0], plane[iy1, ix0], plane[iy0, ix1], plane[iy1, ix1]
            # kernel: column blends, then the sample blend (fp32, FMA where the kernel uses __fmaf_rn)
            cl = fma32(np.array(fy), np.array(p01), np.array(F(ay * p00)))
            cr = fma32(np.array(fy), np.array(p11), np.array(F(ay * p10)))
            ax = F(F(1) - fx)
            a = fma32(np.array(fx), cr, np.array(F(ax * cl)))
            bits = (np.array(a, F).view(np.uint32) & np.uint32(0xffffffc0)) | np.uint32(hi * 8 + wi)
            K.append(bits.view(F))
            # reference: three fp64 products, one all-float product, summed left to right, rounded once
            dfx, dfy = D(fx), D(fy)
            s = (D(1) - dfx) * (D(1) - dfy) * D(p00)
            s = s + (D(1) - dfx) * dfy * D(p01)
            s = s + dfx * (D(1) - dfy) * D(p10)
            s = s + D(F(F(fx * fy) * p11))
            V.append(F(s))
    return np.array(K, F).reshape(-1), np.array(V, F).reshape(-1)


def planes(rng):
    base = rng.standard_normal((30, 30)).astype(F)
    yield "normal", base
    yield "relu", np.maximum(base, 0)
    yield "huge", (base * F(1e30)).astype(F)
    yield "small", (base * F(1e-20)).astype(F)
    yield "offset", (base * F(1e-3) + F(100)).astype(F)          # large common value, small differences
    yield "steps", (F(1) + rng.integers(0, 4, (30, 30)) * F(2.0 ** -23)).astype(F)


def test_key_error_bound_and_decision_rule():
    rng = np.random.default_rng(0)
    worst, decided, total = 0.0, 0, 0
    for name, plane in planes(rng):
        cm = D(np.abs(plane).max())
        for _ in range(60):
            cyx = rng.uniform(0.1, 0.9, 2)
            hw = rng.uniform(16 / 480, 0.9, 2)
            lo = np.clip(cyx - hw / 2, 0, 1).astype(F)
            hi = np.clip(cyx + hw / 2, 0, 1).astype(F)
            roi = [lo[0] + (hi[0] - lo[0]) / F(2), lo[1] + (hi[1] - lo[1]) / F(2), hi[0] - lo[0], hi[1] - lo[1]]
            row, col = int(rng.integers(0, 7)), int(rng.integers(0, 7))
            ys, xs = roi_samples(roi, 30, 30, 7, 7, row, col)
            if len(ys) > 8 or len(xs) > 8:
                continue
            K, V = keys_and_values(plane, ys, xs)
            err = np.abs(K.astype(D) - V.astype(D))
            bound = ALPHA * cm + BETA * np.abs(K.astype(D)) + D(2.0 ** -140)
            worst = max(worst, float((err / bound).max()))
            assert (err <= bound).all(), (name, float((err / bound).max()))
            # the kernel's rule: gap between the two largest keys > T  =>  the exact maximum is unique and is that sample
            order = np.argsort(-K.astype(D), kind="stable")
            total += 1
            if len(K) == 1:
                continue
            kb, k2 = D(K[order[0]]), D(K[order[1]])
            T = D(F(2.0 ** -19) * F(cm)) + D(F(1.375 * 2.0 ** -16)) * abs(kb) + D(2.0 ** -119)
            if kb - k2 > T:
                decided += 1
                b = order[0]
                assert (V[b] > np.delete(V, b)).all(), name
                ref_arg = int(np.argmax(V))          # first maximum, as the reference's strict '<' scan
                assert ref_arg == b
    assert worst <= 1.0 and decided > total // 2    # the bound holds with margin and the fast path is the common case
