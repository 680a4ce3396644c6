"""Deadlock watch for the warp-specialised kernels: several hundred short launches (budget-limited, resumed, re-seeded when a
batch finishes) in a child process with a watchdog.  The round protocol between control and energy warps is only exercised in its
rare orders - very short rounds at the end of a launch, energy warps a round apart - by many launches; an experimental barrier
scheme that passed every functional test deadlocked once in about 140 launches of this loop (profiles/r2_experiments.md)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload", ["h2o20", "sample8"])
def test_many_short_launches_do_not_deadlock(workload):
    env = dict(os.environ, WL=workload)
    try:
        res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "stress_hang.py"), "400", "3000"], env=env, capture_output=True, text=True,
                             timeout=150)
    except subprocess.TimeoutExpired as e:                     # (the child's own watchdog should have fired first)
        pytest.fail(f"stress loop did not finish: {e.stdout}")
    assert res.returncode == 0, res.stdout + res.stderr
    assert "400 launches" in res.stdout and "crew2_kernel" in res.stdout
