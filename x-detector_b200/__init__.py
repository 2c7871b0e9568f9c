"""xdet_b200 -- B200-native Light-Head R-CNN hot path behind the X-Detector API surface.

Host code is Python; all compute is hand-written CUDA for sm_100a in ``csrc/``, reached through
the C-ABI declared in ``include/xdet_b200.h`` (``libxdet_b200.so``, loaded with ctypes).  PyTorch
tensors are only the container for device memory and streams.  There is no CPU fallback: every
operator raises if the native library is missing.
"""
from . import _native  # noqa: F401

__version__ = "0.1"
