"""CPU-only: pin the oracle (oracle/psroi_oracle.c) to the reference.

* against the golden vectors minted by running the reference's own C++ (tests/golden/make_golden.py),
  incl. the known-answer values of the reference fixture quoted in BASELINE.md section 5;
* where /root/reference is mounted (build container), bit-for-bit against the compiled reference
  on fresh random inputs.
"""
import hashlib
import os

import numpy as np
import pytest

from tests import workloads

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "psroi_golden.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


def test_reference_fixture_known_answers(oracle_built):
    # BASELINE.md section 5 / SURVEY.md 8c, minted from the compiled reference
    x, rois = GOLD["fix_inputs"], GOLD["fix_rois"]
    p, i = oracle_built.psroi_align_fwd(x, rois, 2, 2, "mean")
    want = np.array([[5.125, 6.5, 12, 13.375], [9.25, 11.375, 19.875, 22], [17.5, 18.6875, 23.4375, 24.625]],
                    np.float32)
    assert np.array_equal(p[0, :, :, 0], want)
    assert np.array_equal(p, np.repeat(want[None, :, :, None], 4, axis=3))  # 4 identical bank channels
    assert not i.any()
    p, i = oracle_built.psroi_align_fwd(x, rois, 2, 2, "max")
    want = np.array([[7.1875, 8.5625, 14.0625, 15.4375], [13.75, 15.625, 23.125, 25], [19.75, 20.625, 24.125, 25]],
                    np.float32)
    assert np.array_equal(p[0, :, :, 0], want)
    assert np.array_equal(i[0, :, 0, 0], [3, 8, 3])
    for m, total in (("mean", 47.999999), ("max", 48.0)):
        p, i = oracle_built.psroi_align_fwd(x, rois, 2, 2, m)
        g = oracle_built.psroi_align_bwd(x.shape, rois, np.ones_like(p), i, 2, 2, m)
        assert abs(float(g.sum(dtype=np.float64)) - total) < 2e-6


@pytest.mark.parametrize("method", ["mean", "max"])
def test_oracle_matches_golden(oracle_built, method):
    for tag, (x, rois, gw) in {
        "fix": (GOLD["fix_inputs"], GOLD["fix_rois"], 2),
        "small": (workloads.make_map(2, 98, 30, 30, seed=10), workloads.make_rois(2, 28, seed=11, edge_cases=True), 7),
    }.items():
        p, i = oracle_built.psroi_align_fwd(x, rois, gw, gw, method)
        assert np.array_equal(bits(p), bits(GOLD["%s_%s_pooled" % (tag, method)]))
        assert np.array_equal(i, GOLD["%s_%s_index" % (tag, method)])
        gup = np.ones_like(p) if tag == "fix" else np.random.default_rng(12).standard_normal(p.shape, dtype=np.float32)
        g = oracle_built.psroi_align_bwd(x.shape, rois, gup, i, gw, gw, method)
        assert np.array_equal(bits(g), bits(GOLD["%s_%s_grad" % (tag, method)]))


@pytest.mark.parametrize("method", ["mean", "max"])
def test_oracle_config1_digest(oracle_built, method):
    # config 1 (S-model): 1x490x30x30, 300 RoIs + 4 edge cases, 7x7 -- digests minted from the reference
    x = workloads.make_map(1, 490, 30, 30, seed=0)
    rois = workloads.make_rois(1, 300, seed=0, edge_cases=True)
    p, i = oracle_built.psroi_align_fwd(x, rois, 7, 7, method)
    gup = np.random.default_rng(1).standard_normal(p.shape, dtype=np.float32)
    g = oracle_built.psroi_align_bwd(x.shape, rois, gup, i, 7, 7, method)
    assert [sha(p), sha(i), sha(g)] == list(GOLD["c1_%s_sha" % method])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")
@pytest.mark.parametrize("shape", [(2, 490, 30, 30, 100, 7, 7), (1, 900, 30, 30, 64, 15, 15), (1, 98, 50, 37, 50, 7, 7),
                                   (3, 24, 9, 17, 40, 3, 2), (1, 16, 5, 5, 0, 2, 2), (1, 490, 48, 48, 20, 7, 7)])
@pytest.mark.parametrize("method", ["mean", "max"])
def test_oracle_vs_compiled_reference(oracle_built, shape, method):
    N, C, H, W, R, gw, gh = shape
    x = workloads.make_map(N, C, H, W, seed=sum(shape))
    rois = workloads.make_rois(N, R, seed=sum(shape) + 1, min_side=0.02, edge_cases=R > 0)
    p, i = oracle_built.psroi_align_fwd(x, rois, gw, gh, method)
    pr, ir = oracle_built.psroi_align_fwd(x, rois, gw, gh, method, impl="ref")
    ir[workloads.degenerate_mask(rois)] = 0  # the reference leaves those unwritten
    assert np.array_equal(bits(p), bits(pr))
    assert np.array_equal(i, ir)
    gup = np.random.default_rng(7).standard_normal(p.shape, dtype=np.float32)
    g = oracle_built.psroi_align_bwd(x.shape, rois, gup, i, gw, gh, method)
    gr = oracle_built.psroi_align_bwd(x.shape, rois, gup, i, gw, gh, method, impl="ref")
    assert np.array_equal(bits(g), bits(gr))


def test_oracle_threading_is_order_independent(oracle_built):
    x = workloads.make_map(1, 98, 30, 30, seed=3)
    rois = workloads.make_rois(1, 37, seed=4)
    a = oracle_built.psroi_align_fwd(x, rois, 7, 7, "max", threads=1)
    b = oracle_built.psroi_align_fwd(x, rois, 7, 7, "max", threads=5)
    assert np.array_equal(bits(a[0]), bits(b[0])) and np.array_equal(a[1], b[1])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")
@pytest.mark.parametrize("name", sorted(workloads.adversarial_maps()))
def test_oracle_vs_compiled_reference_adversarial_planes(oracle_built, name):
    """The planes the GPU SELECT kernel is attacked with (exact ties, near-ties, zeros, -0.0, huge / subnormal
    magnitudes, NaN / inf): the C restatement must make the reference's choice on every one of them -- first maximum
    wins, strict '<', fp64 blend rounded once -- so that 'kernel == oracle' means 'kernel == reference' there too."""
    x = workloads.adversarial_maps()[name]
    rois = workloads.make_rois(1, 96, seed=123, min_side=0.02, edge_cases=True)
    for method in ("max", "mean"):
        with np.errstate(all="ignore"):
            p, i = oracle_built.psroi_align_fwd(x, rois, 7, 7, method)
            pr, ir = oracle_built.psroi_align_fwd(x, rois, 7, 7, method, impl="ref")
        ir[workloads.degenerate_mask(rois)] = 0
        nan = np.isnan(pr)
        assert np.array_equal(np.isnan(p), nan)
        assert np.array_equal(bits(p)[~nan], bits(pr)[~nan]) and np.array_equal(i, ir)
        gup = np.random.default_rng(8).standard_normal(p.shape, dtype=np.float32)
        with np.errstate(all="ignore"):
            g = oracle_built.psroi_align_bwd(x.shape, rois, gup, i, 7, 7, method)
            gr = oracle_built.psroi_align_bwd(x.shape, rois, gup, i, 7, 7, method, impl="ref")
        assert np.array_equal(bits(g), bits(gr))


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree not mounted")
def test_oracle_vs_compiled_reference_random_geometries(oracle_built):
    """Differential sweep over the op's whole argument space inside its documented contract (boxes within the image):
    ragged maps down to 1x1, non-square grids, banks of 1..4, zero to a dozen RoIs per image, boxes from degenerate
    (zero height / width) to the full image, several images -- C restatement versus the reference's compiled C++,
    bit for bit, forward and backward, both pooling methods."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(st.integers(1, 3), st.integers(1, 5), st.integers(1, 5), st.integers(1, 4), st.integers(1, 40),
           st.integers(1, 40), st.integers(0, 12), st.integers(0, 2 ** 31 - 1), st.sampled_from(["max", "mean"]))
    def run(N, gw, gh, bank, H, W, R, seed, method):
        rng = np.random.default_rng(seed)
        C = gw * gh * bank
        x = rng.standard_normal((N, C, H, W), dtype=np.float32)
        if seed % 3 == 0:
            x = np.maximum(x, 0)                                  # post-ReLU planes: exact ties at zero
        a, b = np.sort(rng.random((N, R, 2, 2), dtype=np.float32), axis=2)[:, :, 0], None
        b = np.sort(rng.random((N, R, 2, 2), dtype=np.float32), axis=2)[:, :, 1]
        lo, hi = np.minimum(a, b), np.maximum(a, b)               # ymin,xmin <= ymax,xmax inside [0,1)
        if R:
            hi[:, 0] = lo[:, 0]                                   # a degenerate box
            if R > 1:
                lo[:, 1], hi[:, 1] = 0.0, 1.0                     # the whole image
        hw = (hi - lo).astype(np.float32)
        rois = np.concatenate([lo + hw / np.float32(2), hw], -1).astype(np.float32)   # (cy, cx, h, w)
        p, i = oracle_built.psroi_align_fwd(x, rois, gw, gh, method)
        pr, ir = oracle_built.psroi_align_fwd(x, rois, gw, gh, method, impl="ref")
        ir[workloads.degenerate_mask(rois)] = 0
        assert np.array_equal(bits(p), bits(pr)) and np.array_equal(i, ir)
        gup = rng.standard_normal(p.shape, dtype=np.float32)
        g = oracle_built.psroi_align_bwd(x.shape, rois, gup, i, gw, gh, method)
        gr = oracle_built.psroi_align_bwd(x.shape, rois, gup, i, gw, gh, method, impl="ref")
        assert np.array_equal(bits(g), bits(gr))

    run()
