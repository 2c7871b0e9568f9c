"""CPU: the staged CUDA kernels (x-detector_b200/csrc/, written without a GPU at hand) executed under a host
emulation -- the kernel's OWN source compiled by g++ against stand-ins for the CUDA built-ins, one real thread per
CUDA thread, driven over the launcher's own grid decomposition (tests/staged/emulate_*.cc).  Proves the index
arithmetic, borders, dilation, channel tail, slab split and shared-memory fold before the first GPU run; says nothing
about performance or about hardware behaviour."""
import glob
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
@pytest.mark.parametrize("src", sorted(glob.glob(os.path.join(HERE, "staged", "emulate_*.cc"))),
                         ids=lambda p: os.path.basename(p))
def test_staged_kernel_under_cpu_emulation(src, tmp_path):
    exe = str(tmp_path / "emulate")
    build = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", exe, src], capture_output=True, text=True,
                           timeout=600)
    assert build.returncode == 0, build.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0 and "emulation ok" in run.stdout and "FAIL" not in run.stdout, run.stdout + run.stderr
