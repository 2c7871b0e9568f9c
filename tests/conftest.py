import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # the product creates missing batch-norm variables with TensorFlow's defaults (identity); the parity tests want
    # NON-trivial scale / shift / moving statistics so that a folding or ordering error cannot hide
    import xdet_b200  # noqa: F401  (import shim -> x-detector_b200/)
    from xdet_b200.net import variables
    variables.RANDOMIZE_BN = True


@pytest.fixture(scope="session")
def oracle_built():
    from oracle import psroi
    psroi.build(ref=os.path.isdir("/root/reference"))
    return psroi
