"""CPU: bench.py's reference arm (the CPU leg the driver runs with --impl reference) prints ONE JSON
line with the contract's keys, and non-zero ranks of a multi-rank launch exit without work."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(extra_env=None, *args):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3", *args],
                          capture_output=True, text=True, env=env, timeout=300)


def test_reference_arm_line_has_the_contract_keys():
    res = _run()
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec_incl_obs" and d["unit"] == "env-steps/s"
    assert d["value"] > 100 and d["vs_baseline"] is None and "workload" in d["config"] and "model" not in d["config"]
    assert "ran" in d["config"]            # the CPU arm says what it actually executed
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    if cb["kind"] == "reference":          # the Python reference itself; the C port rides along as a labelled second number
        assert cb["port"]["kind"] == "port" and cb["port"]["value"] > 1e5
    else:
        assert cb["value"] > 1e5
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    res = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}, "--gpus", "2")
    assert res.returncode == 0 and res.stdout.strip() == ""
