"""bench.py --impl reference: the contract of the JSON line (keys the driver reads), on the smallest workload.  The arm times the
unmodified reference under oracle/_ref on host cores; nothing of the product is loaded."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.ref


def _have_ref():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "bernoulli"))


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref not built")
def test_reference_arm_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "4", "--warmup", "3"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "min_bulk_ESS_per_sec" and j["unit"] == "ESS/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 4 and j["warmup"] == 3
    assert j["value"] > 0 and j["e2e"] == {"value": j["value"], "unit": "ESS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["value"] == j["value"] and cb["cores"] >= 1 and cb["sample"]
    assert j["config"]["workload"].startswith("c1")


@pytest.mark.skipif(not _have_ref(), reason="oracle/_ref not built")
def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c1", "--gpus", "2", "--steps", "4", "--warmup", "3"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""
