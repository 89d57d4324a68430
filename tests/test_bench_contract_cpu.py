"""bench.py's contract on the side that runs without a GPU: the reference arm (`--impl reference`, the reference's own
gridcount() loop from oracle/_ref -- or the oracle port where the reference is not compiled -- on the host cores)
prints ONE JSON line with the keys the driver reads, on the same metric, unit and config as the GPU arm."""
import json
import os
import subprocess
import sys

import gcutil as u

ROOT = u.ROOT


def run_reference_arm(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", *extra],
                         cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                     # exactly one line on stdout
    return json.loads(lines[0])


def test_reference_arm_prints_the_contract_line():
    d = run_reference_arm()
    assert d["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["metric"] == "binned atom-frames/sec" and d["unit"] == "G atom-frames/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert cb["kind"] == ("reference" if u.have_ref() else "port")
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_both_arms_name_the_same_workload():
    """the driver divides the two arms' values: they must describe one workload (same_config)"""
    sys.path.insert(0, ROOT)
    import bench
    ref = run_reference_arm()
    assert ref["config"] == bench.headline_config(1)
