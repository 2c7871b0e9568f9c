"""CPU: the numpy TensorFlow stand-in (oracle/tf_shim) reads the TF 1.x semantics it restates the way the TensorFlow
API documentation's own worked examples do (tests/golden/selftest_tf_shim.py) -- run in a fresh interpreter so that the
stand-in never becomes `tensorflow` inside the test process."""
import os
import subprocess
import sys


def test_stand_in_matches_tf_doc_examples():
    script = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "selftest_tf_shim.py")
    r = subprocess.run([sys.executable, script], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "tf_shim self-test ok" in r.stdout, r.stdout + r.stderr
    assert "tensorflow" not in sys.modules or "tf_shim" not in getattr(sys.modules["tensorflow"], "__file__", "")
