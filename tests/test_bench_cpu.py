"""bench.py's host-side contract, checked without a GPU: the algorithmic FLOP figures of SURVEY.md 8(d), and the
`--impl reference` arm (the unmodified reference model from baseline/_ref on the host cores; oracle port only if unstaged): one JSON line with the agreed keys on
rank 0, silence on the other ranks."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import nvit_oracle as O  # noqa: E402


@pytest.mark.parametrize("name,kohonen,expected", [
    ("tiny", False, 1.5229e9), ("b16", False, 1.4449e11), ("l16", False, 4.9612e11), ("b16", True, 1.5414e11),
])
def test_flops_per_image_match_the_survey(name, kohonen, expected):
    cfg = O.named_config(name, use_kohonen=kohonen)
    got = bench.flops_per_image(cfg, kohonen=kohonen)
    assert abs(got - expected) <= 2e-4 * expected, (got, expected)


def _run_reference(extra_env):
    env = dict(os.environ, **extra_env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny", "--steps", "1",
                        "--warmup", "0", "--gpus", "1"], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_prints_one_json_line_with_the_agreed_keys():
    lines = _run_reference({"RANK": "0", "WORLD_SIZE": "1"})
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["value"] > 0 and d["ms_per_step"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    if os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "nvit", "model.py")):
        assert cb["kind"] == "reference"        # the unmodified reference model is staged: the arm must run it, not the port
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    assert _run_reference({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
