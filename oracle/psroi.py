"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the two CPU PsRoiAlign checkers.

* ``oracle``  : ``oracle/liboracle_psroi.so``  -- this repo's plain-C restatement
                (``oracle/psroi_oracle.c``) of ``cpp/PSROIPooling/ps_roi_align_op.cc:94-193`` and
                ``ps_roi_align_grad_op.cc:200-315``.
* ``ref``     : ``oracle/_ref/libref_psroi.so`` -- the reference's own unmodified C++ op compiled
                against a TensorFlow stand-in (``oracle/ref_shim/tf_shim.h``).  Only present where
                ``make -C oracle ref`` ran with ``/root/reference`` mounted; it travels to the GPU
                box as a prebuilt file.

Both take/return numpy arrays with the operator's shapes:
``inputs[N,C,H,W] f32, rois[N,R,4] f32 (cy,cx,h,w) -> pooled[N,R,G,C/G] f32, index[N,R,G,C/G] i32``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle_psroi.so")
_REF_SO = os.path.join(_HERE, "_ref", "libref_psroi.so")
_F32P = ctypes.POINTER(ctypes.c_float)
_I32P = ctypes.POINTER(ctypes.c_int32)


def build(ref=True):
    """Compile the C restatement (always) and the reference build (when /root/reference is there)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "oracle"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def have_ref():
    return os.path.exists(_REF_SO)


_libs = {}


def _lib(kind):
    if kind not in _libs:
        if kind == "oracle":
            if not os.path.exists(_ORACLE_SO):
                build(ref=False)
            lib = ctypes.CDLL(_ORACLE_SO)
            lib.oracle_psroi_align_fwd.argtypes = [_F32P, _F32P, _F32P, _I32P] + [ctypes.c_int] * 9
            lib.oracle_psroi_align_bwd.argtypes = [_F32P, _F32P, _I32P, _F32P] + [ctypes.c_int] * 9
        else:
            if not have_ref():
                raise FileNotFoundError(_REF_SO + " (run `make -C oracle ref` where /root/reference exists)")
            lib = ctypes.CDLL(_REF_SO)
            lib.ref_psroi_align_fwd.argtypes = [_F32P, _F32P, _F32P, _I32P] + [ctypes.c_int] * 9
            lib.ref_psroi_align_bwd.argtypes = [_F32P, _F32P, _F32P, _I32P, _F32P] + [ctypes.c_int] * 9
            lib.ref_psroi_last_error_fwd.restype = ctypes.c_char_p
            lib.ref_psroi_last_error_bwd.restype = ctypes.c_char_p
        _libs[kind] = lib
    return _libs[kind]


def _f(a):
    return a.ctypes.data_as(_F32P)


def _i(a):
    return a.ctypes.data_as(_I32P)


def default_threads():
    return os.cpu_count() or 1


def psroi_align_fwd(inputs, rois, gw, gh, pool_method="max", threads=None, impl="oracle", index_fill=0):
    """CPU PsRoiAlign forward.  impl = "oracle" (C restatement) | "ref" (compiled reference)."""
    inputs = np.ascontiguousarray(inputs, dtype=np.float32)
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    N, C, H, W = inputs.shape
    R = rois.shape[1]
    G = gw * gh
    use_max = 1 if "max" in pool_method else 0
    threads = threads or default_threads()
    out = np.zeros((N, R, G, C // G), np.float32)
    idx = np.full((N, R, G, C // G), index_fill, np.int32)  # the reference leaves degenerate entries unwritten
    if impl == "oracle":
        rc = _lib("oracle").oracle_psroi_align_fwd(_f(inputs), _f(rois), _f(out), _i(idx), N, C, H, W, R, gw, gh,
                                                   use_max, threads)
    else:
        rc = _lib("ref").ref_psroi_align_fwd(_f(inputs), _f(rois), _f(out), _i(idx), N, C, H, W, R, gw, gh,
                                             use_max, threads)
    if rc != 0:
        msg = _lib("ref").ref_psroi_last_error_fwd().decode() if impl == "ref" else "bad arguments"
        raise ValueError("psroi_align_fwd[%s] failed rc=%d %s" % (impl, rc, msg))
    return out, idx


def psroi_align_bwd(inputs_shape, rois, pooled_grad, index, gw, gh, pool_method="max", threads=None, impl="oracle"):
    """CPU PsRoiAlignGrad.  Returns grad[N,C,H,W]."""
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    pooled_grad = np.ascontiguousarray(pooled_grad, dtype=np.float32)
    index = np.ascontiguousarray(index, dtype=np.int32)
    N, C, H, W = inputs_shape
    R = rois.shape[1]
    use_max = 1 if "max" in pool_method else 0
    threads = threads or default_threads()
    grad = np.empty((N, C, H, W), np.float32)
    if impl == "oracle":
        rc = _lib("oracle").oracle_psroi_align_bwd(_f(rois), _f(pooled_grad), _i(index), _f(grad), N, C, H, W, R, gw,
                                                   gh, use_max, threads)
    else:
        dummy = np.zeros((N, C, H, W), np.float32)  # the reference op only reads its shape
        rc = _lib("ref").ref_psroi_align_bwd(_f(dummy), _f(rois), _f(pooled_grad), _i(index), _f(grad), N, C, H, W, R,
                                             gw, gh, use_max, threads)
    if rc != 0:
        msg = _lib("ref").ref_psroi_last_error_bwd().decode() if impl == "ref" else "bad arguments"
        raise ValueError("psroi_align_bwd[%s] failed rc=%d %s" % (impl, rc, msg))
    return grad
