"""GPU parity tests: the CUDA PsRoiAlign (through the C-ABI) against the CPU oracle and the
golden vectors minted from the reference.  Bar: bit-exact features, indices and gradients."""
import hashlib
import os

import numpy as np
import pytest

from tests import workloads

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "psroi_golden.npz"))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return np.ascontiguousarray(a).view(np.int32)


@pytest.fixture(scope="module")
def ops():
    import torch
    assert torch.cuda.is_available()
    import xdet_b200  # noqa: F401
    from xdet_b200 import _native, ops
    _native.lib()  # must exist: no fallback
    return ops


def run_fwd(ops, x, rois, gw, gh, method, variant="auto"):
    import torch
    p, i = ops.ps_roi_align(torch.from_numpy(x).cuda(), torch.from_numpy(rois).cuda(), gw, gh, method, variant=variant)
    torch.cuda.synchronize()
    return p.cpu().numpy(), i.cpu().numpy()


def run_bwd(ops, shape, rois, g, idx, gw, gh, method):
    import torch
    x = torch.empty(shape, device="cuda")
    out = ops.ps_roi_align_grad(x, torch.from_numpy(rois).cuda(), torch.from_numpy(g).cuda(),
                                torch.from_numpy(idx).cuda(), gw, gh, method)
    torch.cuda.synchronize()
    return out.cpu().numpy()


def variants_for(method):
    """SELECT (fp32 selection + exact fp64 blend of the selected sample) exists for max pooling only."""
    return ("gather", "planes", "select") if method == "max" else ("gather", "planes")


@pytest.mark.parametrize("variant", ["gather", "planes", "select", "auto"])
@pytest.mark.parametrize("method", ["mean", "max"])
def test_reference_fixture(ops, method, variant):
    if variant == "select" and method != "max":
        pytest.skip("SELECT is max-only")
    x, rois = GOLD["fix_inputs"], GOLD["fix_rois"]
    p, i = run_fwd(ops, x, rois, 2, 2, method, variant)
    assert np.array_equal(bits(p), bits(GOLD["fix_%s_pooled" % method]))
    assert np.array_equal(i, GOLD["fix_%s_index" % method])
    g = run_bwd(ops, x.shape, rois, np.ones_like(p), i, 2, 2, method)
    assert np.array_equal(bits(g), bits(GOLD["fix_%s_grad" % method]))


@pytest.mark.parametrize("variant", ["gather", "planes", "select"])
@pytest.mark.parametrize("method", ["mean", "max"])
def test_small_golden(ops, method, variant):
    if variant == "select" and method != "max":
        pytest.skip("SELECT is max-only")
    x = workloads.make_map(2, 98, 30, 30, seed=10)
    rois = workloads.make_rois(2, 28, seed=11, edge_cases=True)
    p, i = run_fwd(ops, x, rois, 7, 7, method, variant)
    assert np.array_equal(bits(p), bits(GOLD["small_%s_pooled" % method]))
    assert np.array_equal(i, GOLD["small_%s_index" % method])
    gup = np.random.default_rng(12).standard_normal(p.shape, dtype=np.float32)
    g = run_bwd(ops, x.shape, rois, gup, i, 7, 7, method)
    assert np.array_equal(bits(g), bits(GOLD["small_%s_grad" % method]))


@pytest.mark.parametrize("variant", ["gather", "planes", "select"])
@pytest.mark.parametrize("method", ["mean", "max"])
def test_config1_model_shape_digest(ops, method, variant):
    """BASELINE config 1 (S-model): 1x490x30x30 + 300 RoIs (+4 edge cases), 7x7; pinned to digests
    of the reference's own output."""
    if variant == "select" and method != "max":
        pytest.skip("SELECT is max-only")
    x = workloads.make_map(1, 490, 30, 30, seed=0)
    rois = workloads.make_rois(1, 300, seed=0, edge_cases=True)
    p, i = run_fwd(ops, x, rois, 7, 7, method, variant)
    gup = np.random.default_rng(1).standard_normal(p.shape, dtype=np.float32)
    g = run_bwd(ops, x.shape, rois, gup, i, 7, 7, method)
    assert [sha(p), sha(i), sha(g)] == list(GOLD["c1_%s_sha" % method])


SHAPES = [
    (2, 490, 30, 30, 300, 7, 7),     # model shape, batch 2
    (1, 980, 30, 30, 512, 7, 7),     # sweep channel count, bank 20
    (1, 900, 30, 30, 256, 15, 15),   # 15x15 bins, bank 4
    (1, 490, 50, 50, 200, 7, 7),     # 800x800 image map
    (3, 24, 9, 17, 40, 3, 2),        # ragged: gw != gh, H != W
    (1, 98, 50, 37, 33, 7, 7),
    (1, 16, 5, 5, 0, 2, 2),          # empty RoI set
    (1, 49, 30, 30, 1, 7, 7),        # single RoI, bank 1
    (1, 98, 12, 12, 1100, 7, 7),     # > 1024 RoIs: the backward's thread-per-plane kernel (<= 1024: warp per plane)
    (8, 490, 30, 30, 64, 7, 7),      # the training step's shape (8 images x 64 sampled RoIs)
]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("method", ["mean", "max"])
def test_fwd_bwd_vs_oracle(ops, oracle_built, shape, method):
    N, C, H, W, R, gw, gh = shape
    x = workloads.make_map(N, C, H, W, seed=sum(shape))
    rois = workloads.make_rois(N, R, seed=sum(shape) + 1, min_side=0.02, edge_cases=R > 0)
    po, io = oracle_built.psroi_align_fwd(x, rois, gw, gh, method)
    for variant in variants_for(method):
        p, i = run_fwd(ops, x, rois, gw, gh, method, variant)
        assert np.array_equal(bits(p), bits(po)), variant
        assert np.array_equal(i, io), variant
    gup = np.random.default_rng(7).standard_normal(po.shape, dtype=np.float32)
    g = run_bwd(ops, x.shape, rois, gup, io, gw, gh, method)
    go = oracle_built.psroi_align_bwd(x.shape, rois, gup, io, gw, gh, method)
    assert np.array_equal(bits(g), bits(go))


_adversarial_maps = workloads.adversarial_maps


@pytest.mark.parametrize("name", sorted(_adversarial_maps()))
def test_select_adversarial_planes(ops, oracle_built, name):
    x = _adversarial_maps()[name]
    rois = workloads.make_rois(1, 96, seed=123, min_side=0.02, edge_cases=True)
    with np.errstate(all="ignore"):
        po, io = oracle_built.psroi_align_fwd(x, rois, 7, 7, "max")
    nan = np.isnan(po)   # NaN payloads differ between x86 and the GPU (0xFFC00000 vs 0x7FFFFFFF): compare positions
    for variant in ("select", "planes", "gather"):
        p, i = run_fwd(ops, x, rois, 7, 7, "max", variant)
        assert np.array_equal(np.isnan(p), nan), (name, variant)
        assert np.array_equal(bits(p)[~nan], bits(po)[~nan]), (name, variant)
        assert np.array_equal(i, io), (name, variant)


def test_select_many_samples_per_bin(ops, oracle_built):
    """Big RoIs on a 50x50 map with 2x2 bins: up to 26 samples per bin axis, beyond the 8-deep sample tables
    (generic path inside SELECT), mixed with small RoIs that use the tables."""
    x = workloads.make_map(1, 64, 50, 50, seed=31)
    rois = workloads.make_rois(1, 64, seed=32, min_side=0.02, max_side=1.0, edge_cases=True)
    po, io = oracle_built.psroi_align_fwd(x, rois, 2, 2, "max")
    p, i = run_fwd(ops, x, rois, 2, 2, "max", "select")
    assert np.array_equal(bits(p), bits(po)) and np.array_equal(i, io)


def test_config1_literal_480_map(ops, oracle_built):
    """BASELINE config 1 read literally: a 1x490x480x480 map (451 MB >> L2), 300 RoIs, 7x7 max.
    The planes do not fit shared memory, so this exercises the global-gather kernel."""
    x = workloads.make_map(1, 490, 480, 480, seed=0)
    rois = workloads.make_rois(1, 300, seed=0, edge_cases=True)
    po, io = oracle_built.psroi_align_fwd(x, rois, 7, 7, "max")
    p, i = run_fwd(ops, x, rois, 7, 7, "max", "auto")
    assert np.array_equal(bits(p), bits(po)) and np.array_equal(i, io)


def test_sweep_top_properties(ops):
    """Full sweep size (R=16384, C=980, 7x7; BASELINE config 5's top point, 16 M outputs): every kernel variant
    against the ORACLE (all host threads), bit for bit -- features and arg-max indices, max and mean -- plus the
    size-independent properties: a RoI's result does not depend on its neighbours (permutation equivariance),
    mean <= max, indices in range."""
    import os

    import torch
    from oracle import psroi as oracle_psroi
    x = workloads.make_map(1, 980, 30, 30, seed=4)
    rois = workloads.make_rois(1, 16384, seed=5)
    pa, ia = run_fwd(ops, x, rois, 7, 7, "max", "planes")
    po, io = oracle_psroi.psroi_align_fwd(x, rois, 7, 7, "max", threads=os.cpu_count() or 1)
    assert np.array_equal(bits(pa), bits(po)) and np.array_equal(ia, io)
    pmo, _ = oracle_psroi.psroi_align_fwd(x, rois, 7, 7, "mean", threads=os.cpu_count() or 1)
    pb, ib = run_fwd(ops, x, rois, 7, 7, "max", "gather")
    assert np.array_equal(bits(pa), bits(pb)) and np.array_equal(ia, ib)
    ps, is_ = run_fwd(ops, x, rois, 7, 7, "max", "select")   # 16 M outputs: the near-tie fallback is exercised
    assert np.array_equal(bits(ps), bits(pa)) and np.array_equal(is_, ia)
    perm = np.random.default_rng(6).permutation(16384)
    pp, ip = run_fwd(ops, x, np.ascontiguousarray(rois[:, perm]), 7, 7, "max", "planes")
    assert np.array_equal(bits(pp), bits(pa[:, perm])) and np.array_equal(ip, ia[:, perm])
    pm, im = run_fwd(ops, x, rois, 7, 7, "mean", "planes")
    assert np.array_equal(bits(pm), bits(pmo))
    assert (pm <= pa + 1e-6).all() and not im.any()
    assert ia.min() >= 0 and ia.max() < 25  # at most 5x5 samples per bin on a 30x30 map with 7x7 bins
    del torch


def test_autograd_pairing(ops):
    import torch
    x = torch.from_numpy(workloads.make_map(1, 98, 30, 30, seed=1)).cuda().requires_grad_(True)
    rois = torch.from_numpy(workloads.make_rois(1, 16, seed=2)).cuda()
    pooled, index = ops.PsRoiAlign.apply(x, rois, 7, 7, "max")
    pooled.sum().backward()
    ref = ops.ps_roi_align_grad(x.detach(), rois, torch.ones_like(pooled), index, 7, 7, "max")
    assert torch.equal(x.grad, ref)
    # all-ones upstream: every (roi, channel) distributes exactly weight 1 over its 4 taps
    assert abs(float(x.grad.double().sum()) - pooled.numel()) < 1e-2


def test_host_buffer_entry_point(ops, oracle_built):
    import ctypes
    from xdet_b200 import _native
    x = workloads.make_map(1, 98, 30, 30, seed=8)
    rois = workloads.make_rois(1, 20, seed=9)
    p = np.empty((1, 20, 49, 2), np.float32)
    i = np.empty((1, 20, 49, 2), np.int32)
    rc = _native.lib().xdet_psroi_align_fwd_host(x.ctypes.data_as(ctypes.c_void_p), rois.ctypes.data_as(ctypes.c_void_p),
                                                 p.ctypes.data_as(ctypes.c_void_p), i.ctypes.data_as(ctypes.c_void_p),
                                                 1, 98, 30, 30, 20, 7, 7, 1)
    _native.check(rc)
    po, io = oracle_built.psroi_align_fwd(x, rois, 7, 7, "max")
    assert np.array_equal(bits(p), bits(po)) and np.array_equal(i, io)
