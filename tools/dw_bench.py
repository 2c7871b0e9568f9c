#!/usr/bin/env python
"""Times the fp32 depthwise 3x3 kernel of the f16x2 precision (both implementations, every output form) on the Xception
middle-flow shapes of config 3 (batch 32, 800^2 -> 50x50x728) and config 2-sized maps."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch  # noqa: E402

import xdet_b200  # noqa: E402,F401
from xdet_b200 import _native, ops  # noqa: E402
from xdet_b200.ops import conv as conv_ops  # noqa: E402


def timed(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


def main():
    for shape in ((32, 50, 50, 728), (8, 30, 30, 728), (32, 100, 100, 256), (32, 200, 200, 128)):
        x = torch.randn(shape, device="cuda")
        w9 = torch.randn((9, shape[-1]), device="cuda")
        nb = x.numel() * 4
        for tma in (1, 0):
            _native.lib().xdet_set_depthwise_f32_tma(tma)
            with conv_ops.precision("f16x2"):
                for forms in ("pair", "f32", "both"):
                    t = timed(lambda: ops.depthwise3x3(x, w9, dilation=1, relu_in=True, forms=forms))
                    wb = nb * {"pair": 1, "f32": 1, "both": 2}[forms]
                    print("%-20s tma=%d forms=%-4s %8.1f us  %6.0f GB/s (read + written)" % (
                        "x".join(map(str, shape)), tma, forms, t, (nb + wb) / t / 1e3))
        _native.lib().xdet_set_depthwise_f32_tma(1)


if __name__ == "__main__":
    main()
