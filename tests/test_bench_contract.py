"""CPU: the reference arm of bench.py (`--impl reference`) runs without a GPU and prints ONE JSON line with the keys the
driver reads; the own arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e, capture_output=True,
                          text=True, timeout=600)


def test_reference_arm_psroi_line():
    r = _run(["--impl", "reference", "--workload", "psroi_sweep_top", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "psroi_align_fwd_GBps" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("psroi_align_fwd")


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(["--impl", "reference", "--workload", "psroi_sweep_top", "--steps", "1", "--warmup", "0", "--gpus", "2"],
             env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""
