"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/*.h declares (no compute calls here -- there is no GPU in the build container)."""
import ctypes
import glob
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    import xdet_b200  # noqa: F401
    from xdet_b200 import _native
    _native.build()
    return _native


def declared_symbols():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names += re.findall(r"\b(xdet_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_something():
    syms = declared_symbols()
    assert "xdet_psroi_align_fwd" in syms and "xdet_psroi_align_bwd" in syms


def test_library_exports_every_declared_symbol(native):
    lib = ctypes.CDLL(native.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_version_and_error_string(native):
    lib = native.lib()
    assert b"sm_100a" in lib.xdet_version()
    assert lib.xdet_last_error() is not None


def test_argument_validation_needs_no_gpu(native):
    # shape/attr violations are rejected before any CUDA call (PSROIAlignOp checks, ps_roi_align_op.cc:208-226)
    lib = native.lib()
    rc = lib.xdet_psroi_align_fwd(None, None, None, None, 1, 490, 30, 30, 10, 0, 7, 1, None)
    assert rc == -1 and b"grid_dim" in lib.xdet_last_error()
    rc = lib.xdet_psroi_align_fwd(None, None, None, None, 1, 1024, 30, 30, 10, 7, 7, 1, None)
    assert rc == -1 and b"divisible" in lib.xdet_last_error()


def test_operator_wrapper_validation(native):
    import torch
    from xdet_b200.ops import ps_roi_align
    x = torch.zeros(1, 490, 30, 30)
    r = torch.zeros(1, 5, 4)
    with pytest.raises(ValueError, match="pool_method"):
        ps_roi_align(x, r, 7, 7, "avg")
    with pytest.raises(ValueError, match="NCHW"):
        ps_roi_align(x[0], r, 7, 7, "max")
    with pytest.raises(ValueError, match="batch_size x num_rois x 4"):
        ps_roi_align(x, r[0], 7, 7, "max")
    with pytest.raises(ValueError, match="don't match"):
        ps_roi_align(x, torch.zeros(2, 5, 4), 7, 7, "max")
    with pytest.raises(ValueError, match="GPU only"):
        ps_roi_align(x, r, 7, 7, "max")  # CPU tensors: no fallback


def test_product_fails_loudly_without_the_cuda_library(tmp_path):
    """No CPU / PyTorch fallback exists: with the shared library absent an op raises NativeLibraryMissing (the
    message says how to build it), and ops that validate their arguments first refuse CPU tensors outright -- checked in a fresh interpreter through an op that needs no GPU to reach the
    library lookup."""
    import os
    import subprocess
    import sys
    code = (
        "import sys, torch\n"
        "sys.path.insert(0, %r)\n"
        "import xdet_b200\n"
        "from xdet_b200 import _native, ops\n"
        "assert not __import__('os').path.exists(_native.LIB_PATH)\n"
        "for call in (lambda: _native.lib(),\n"
        "             lambda: ops.affine_relu(torch.zeros((1, 2, 2, 8), dtype=torch.bfloat16), torch.ones(8), torch.zeros(8)),\n"
        "             lambda: ops.ps_roi_align(torch.zeros((1, 4, 5, 5)), torch.zeros((1, 1, 4)), 2, 2, 'max')):\n"
        "    try:\n"
        "        call()\n"
        "    except (_native.NativeLibraryMissing, ValueError) as e:\n"
        "        assert 'no CPU or PyTorch fallback' in str(e) or 'no CPU fallback' in str(e), e\n"
        "    else:\n"
        "        raise SystemExit('an op ran without the CUDA library')\n"
        "print('LOUD')\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, XDET_B200_LIB=str(tmp_path / "absent" / "libxdet_b200.so"))
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "LOUD" in r.stdout, r.stdout + r.stderr


def test_ctypes_structures_match_the_header(tmp_path):
    """The Python side fills the descriptor structs of include/xdet_b200.h through ctypes mirrors: size and the offset of
    every field must equal what the C compiler lays out (a silent drift would corrupt every launch)."""
    import subprocess
    import xdet_b200  # noqa: F401
    from xdet_b200.ops import conv, train
    pairs = [("xdet_conv_desc", conv.ConvDesc), ("xdet_conv_f16x2_desc", conv.ConvF16x2Desc),
             ("xdet_wgrad_desc", conv.WgradDesc), ("xdet_sgd_item", train.SgdItem)]
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "xdet_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append('  printf("%s %%zu", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf(" %%zu", offsetof(%s, %s));' % (cname, fname))
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    assert len(out) == len(pairs)
    for line, (cname, cls) in zip(out, pairs):
        tok = line.split()
        assert tok[0] == cname
        assert int(tok[1]) == ctypes.sizeof(cls), (cname, tok[1], ctypes.sizeof(cls))
        offs = [int(t) for t in tok[2:]]
        assert offs == [getattr(cls, f).offset for f, _ in cls._fields_], cname
