"""bench.py's output contract on the CPU-runnable arm (`--impl reference`: the fast CPU port of the
reference's path, oracle/qn_fastport.c): stdout carries exactly ONE line, that line is JSON, and it
has the keys the driver reads.  (The GPU arm prints through the same `emit`.)"""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(extra_env=None):
    env = dict(os.environ)
    env.update(extra_env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-sample", "64"], capture_output=True, text=True, cwd=ROOT, env=env,
                       timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line():
    out = _run()
    lines = [l for l in out.split("\n") if l.strip()]
    assert len(lines) == 1, out[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("converged BFGS solves/sec (n=60 Rosenbrock, 2^16 batch)")
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]


def test_reference_arm_other_ranks_print_nothing():
    """Under torchrun (N > 1) rank 0 alone runs the CPU arm; the others exit 0 without output."""
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.strip() == ""
