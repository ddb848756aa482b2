"""The reference arm of bench.py (CPU only): runs on a small sector and emits the contract's keys.
The arm's value must be the reference's DEFAULT stored-table path measured on full products; the
non-default direct path is reported as an explicitly extrapolated side figure."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--ns", "10",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "hxv_per_s" and d["unit"] == "Hxv/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] - 1e3) < 1e-6 * 1e3
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and "FULL products" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Hxv/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["direct_variant"]["extrapolated"] is True and d["direct_variant"]["value"] > 0
    assert d["config"]["ns"] == 10 and "workload" in d["config"]


def test_reference_arm_nonzero_rank_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--ns", "8"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
