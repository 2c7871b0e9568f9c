"""Loader / builder for the native library (``csrc/libxdet_b200.so``) and ctypes prototypes of
the C-ABI in ``include/xdet_b200.h``."""
import ctypes
import glob
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# XDET_B200_LIB: load another build of the same library (A/B experiments of tools/); default = the in-tree build
LIB_PATH = os.environ.get("XDET_B200_LIB") or os.path.join(CSRC, "libxdet_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    # bit-exact parity with the reference's CPU arithmetic needs separate mul/add roundings
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-shared",
]


class NativeLibraryMissing(RuntimeError):
    pass


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def build(force=False, verbose=False):
    """Compile every ``csrc/*.cu`` for sm_100a into ``csrc/libxdet_b200.so`` (in-tree)."""
    srcs = sources()
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    if not force and os.path.exists(LIB_PATH):
        if os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(p) for p in deps):
            return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + srcs + ["-o", LIB_PATH]
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_PATH


_lib = None
_F32P = ctypes.POINTER(ctypes.c_float)
_I32P = ctypes.POINTER(ctypes.c_int32)


def _declare(lib):
    c_int, c_void_p = ctypes.c_int, ctypes.c_void_p
    lib.xdet_last_error.restype = ctypes.c_char_p
    lib.xdet_version.restype = ctypes.c_char_p
    lib.xdet_launch_count.restype = ctypes.c_longlong
    lib.xdet_psroi_align_fwd.argtypes = [c_void_p] * 4 + [c_int] * 8 + [c_void_p]
    lib.xdet_psroi_align_fwd_ex.argtypes = [c_void_p] * 4 + [c_int] * 9 + [c_void_p]
    lib.xdet_psroi_align_bwd.argtypes = [c_void_p] * 4 + [c_int] * 8 + [c_void_p]
    lib.xdet_psroi_align_fwd_host.argtypes = [c_void_p] * 4 + [c_int] * 8
    lib.xdet_psroi_align_bwd_host.argtypes = [c_void_p] * 4 + [c_int] * 8
    lib.xdet_conv2d_bf16.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.xdet_conv2d_wgrad_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p]
    lib.xdet_conv2d_wgrad_f16x2.argtypes = [c_void_p, ctypes.c_longlong, c_void_p, ctypes.c_longlong, c_void_p, ctypes.c_float,
                                            c_void_p]
    c_ll, c_float, c_size_t = ctypes.c_longlong, ctypes.c_float, ctypes.c_size_t
    lib.xdet_im2col_bf16.argtypes = [c_void_p, c_int, c_void_p] + [c_int] * 13 + [c_void_p]
    lib.xdet_maxpool3x3s2_bf16.argtypes = [c_void_p] * 5 + [c_int] * 8 + [c_void_p]
    lib.xdet_maxpool3x3s2_add_bf16.argtypes = [c_void_p] * 6 + [c_int] * 8 + [c_void_p]
    lib.xdet_maxpool3x3s2_argmax_bf16.argtypes = [c_void_p] * 3 + [c_int] * 8 + [c_void_p]
    lib.xdet_depthwise3x3_bf16.argtypes = [c_void_p] * 3 + [c_int] * 6 + [c_void_p]
    lib.xdet_affine_relu_bf16.argtypes = [c_void_p] * 4 + [c_ll, c_int, c_int, c_void_p]
    lib.xdet_f32_to_bf16_rows.argtypes = [c_void_p, c_void_p, c_ll, c_int, c_int, c_void_p]
    lib.xdet_image_to_nhwc8_bf16.argtypes = [c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]
    lib.xdet_det_postprocess_workspace_bytes.argtypes = [c_int] * 4
    lib.xdet_det_postprocess_workspace_bytes.restype = c_size_t
    lib.xdet_det_postprocess.argtypes = ([c_void_p] * 4 + [c_int] * 3 + [c_float, c_int, c_int, c_float] +
                                         [c_void_p] * 3 + [c_size_t, c_void_p])
    lib.xdet_set_depthwise_rows.argtypes = [c_int]
    lib.xdet_preprocess_eval_u8.argtypes = [c_void_p, c_int, c_int, c_int, c_int, ctypes.POINTER(ctypes.c_float * 3),
                                            c_void_p, c_void_p]
    lib.xdet_det_match.argtypes = [c_void_p] * 4 + [c_int] * 4 + [c_float] + [c_void_p] * 4
    lib.xdet_split3_bf16.argtypes = [c_void_p] + [c_ll] * 4 + [c_int] * 4 + [c_void_p, c_int, c_void_p]
    lib.xdet_conv2d_f16x2.argtypes = [c_void_p, c_void_p, c_void_p]
    lib.xdet_split2_f16.argtypes = [c_void_p] + [c_ll] * 4 + [c_int] * 4 + [c_void_p, c_int, c_int, c_int, c_ll, c_int,
                                                                            c_void_p]
    lib.xdet_maxpool3x3s2_f32x.argtypes = [c_void_p] * 8 + [c_ll] + [c_int] * 8 + [c_void_p]
    lib.xdet_depthwise3x3_f32x.argtypes = [c_void_p] * 4 + [c_ll] + [c_int] * 6 + [c_void_p]
    lib.xdet_f32_post.argtypes = [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_ll, c_int,
                                  c_void_p]
    lib.xdet_maxpool3x3s2_f32.argtypes = [c_void_p] * 6 + [c_int] * 8 + [c_void_p]
    lib.xdet_depthwise3x3_f32.argtypes = [c_void_p] * 3 + [c_int] * 6 + [c_void_p]
    lib.xdet_rpn_decode.argtypes = [c_void_p, c_int, c_int, c_int] + [c_void_p] * 4 + [c_int] * 4 + [c_void_p] * 3
    lib.xdet_head_decode.argtypes = [c_void_p, c_void_p] + [c_int] * 4 + [c_ll, c_void_p, c_void_p, c_void_p]
    lib.xdet_head_decode_ex.argtypes = [c_void_p, c_void_p] + [c_int] * 4 + [c_ll] + [c_void_p] * 5
    lib.xdet_rpn_select_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.xdet_rpn_select_workspace_bytes.restype = c_size_t
    lib.xdet_rpn_select.argtypes = ([c_void_p, c_void_p] + [c_int] * 4 + [c_float, c_float] + [c_void_p] * 6 +
                                    [c_size_t, c_void_p])


def _declare_train(lib):
    c_int, c_void_p, c_ll, c_float, c_size_t = (ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_float,
                                                 ctypes.c_size_t)
    lib.xdet_col_stats_bf16.argtypes = [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.xdet_bn_finalize.argtypes = [c_void_p] * 3 + [c_ll, c_int, c_float, c_float] + [c_void_p] * 7
    lib.xdet_bn_relu_bwd_bf16.argtypes = [c_void_p] * 6 + [c_ll, c_int, c_int] + [c_void_p] * 3 + [c_int, c_void_p]
    lib.xdet_bn_train_apply_bf16.argtypes = [c_void_p, c_void_p, c_ll, c_int, c_void_p, c_void_p, c_void_p, c_float,
                                             c_float] + [c_void_p] * 6 + [c_int, c_void_p]
    lib.xdet_bn_train_scratch_bytes.restype = c_size_t
    lib.xdet_bn_train_scratch_bytes.argtypes = [c_int]
    lib.xdet_bn_train_stats_bf16.argtypes = [c_void_p, c_ll, c_int, c_int, c_void_p, c_void_p, c_float, c_float] + \
        [c_void_p] * 8
    lib.xdet_depthwise3x3_wgrad_bf16.argtypes = [c_void_p] * 3 + [c_int] * 6 + [c_void_p]
    lib.xdet_col_stats_f32.argtypes = [c_void_p, c_ll, c_int, c_int, c_int, c_int, c_void_p, c_void_p]
    lib.xdet_bn_relu_bwd_f32.argtypes = [c_void_p] * 6 + [c_ll, c_int, c_int] + [c_void_p] * 4
    lib.xdet_depthwise3x3_wgrad_f32.argtypes = [c_void_p] * 3 + [c_int] * 6 + [c_void_p]
    lib.xdet_relu_bwd_f32.argtypes = [c_void_p, c_void_p, c_void_p, c_ll, c_void_p]
    lib.xdet_maxpool3x3s2_argmax_f32.argtypes = [c_void_p] * 3 + [c_int] * 8 + [c_void_p]
    lib.xdet_maxpool3x3s2_bwd_f32.argtypes = [c_void_p] * 3 + [c_int] * 8 + [c_void_p]
    lib.xdet_relu_bwd_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_ll, c_void_p]
    lib.xdet_maxpool3x3s2_bwd_bf16.argtypes = [c_void_p] * 3 + [c_int] * 8 + [c_void_p]
    lib.xdet_nchw_f32_to_nhwc_bf16.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    lib.xdet_affine_relu_to_nchw_f32.argtypes = [c_void_p] * 4 + [c_int] * 5 + [c_void_p]
    lib.xdet_softmax_ce.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_float, c_ll, c_void_p, c_void_p, c_int,
                                    c_void_p]
    lib.xdet_smooth_l1.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_float, c_ll, c_void_p, c_void_p, c_int, c_void_p]
    lib.xdet_sgd_momentum_conv.argtypes = [c_void_p] * 5 + [c_int] * 9 + [c_float] * 4 + [c_void_p]
    lib.xdet_sgd_momentum_vec.argtypes = [c_void_p] * 3 + [c_ll] + [c_float] * 4 + [c_void_p]
    lib.xdet_set_depthwise_f32_tma.argtypes = [c_int]
    lib.xdet_set_depthwise_f32_tma.restype = None
    lib.xdet_sgd_momentum_multi.argtypes = [c_void_p, c_int, c_int, c_float, c_float, c_float, c_void_p]
    lib.xdet_match_workspace_bytes.argtypes = [c_int, c_int]
    lib.xdet_match_workspace_bytes.restype = c_size_t
    lib.xdet_match_encode.argtypes = ([c_void_p, c_ll] + [c_void_p] * 3 + [c_int] * 3 + [c_float] * 3 +
                                      [ctypes.POINTER(c_float)] + [c_void_p] * 5)
    lib.xdet_sample_fg_bg.argtypes = [c_void_p, c_void_p, c_float] + [c_int] * 4 + [c_void_p] * 7


def lib():
    """The loaded C-ABI library.  Raises loudly if it has not been built: no fallback exists."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeLibraryMissing(
                "%s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). xdet_b200 has no CPU or PyTorch fallback." % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _declare(_lib)
        _declare_train(_lib)
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().xdet_last_error().decode()
        if rc == -1:
            raise ValueError(msg)
        raise RuntimeError("xdet_b200 native error %d: %s" % (rc, msg))


def launch_count():
    return int(lib().xdet_launch_count())
