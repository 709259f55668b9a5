"""bench.py's reference arm on the host (no GPU): the JSON contract of `--impl reference`, that it runs the UNMODIFIED
reference modules when they are present (kind "reference") and says which sample it timed, and that the files staged for the
GPU box under oracle/_ref/reference are byte-identical to the hashes committed in oracle/reference.sha256."""
import hashlib
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    env = dict(os.environ, DN_BENCH_CPU_SAMPLE="1x48")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0 and line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["cores"] >= 1 and cb["kind"] in ("reference", "port")
    from oracle import ref_loader
    assert cb["kind"] == ("reference" if ref_loader.available() else "port")
    # the arm states the sample it ran and that the loop is extrapolated — in the workload string and as fields
    assert "B 1 x T 48" in line["config"]["workload"] and "extrapolated" in line["config"]["workload"]
    assert line["config"]["sample"] == {"batch": 1, "frames": 48, "calls_timed": 2, "calls_extrapolated_to": 99}


def test_staged_reference_files_match_committed_hashes():
    staged = os.path.join(ROOT, "oracle", "_ref", "reference")
    if not os.path.isdir(staged):
        pytest.skip("oracle/_ref/reference not staged (make -C oracle ref)")
    want = {}
    for ln in open(os.path.join(ROOT, "oracle", "reference.sha256")):
        h, name = ln.split()
        want[name] = h
    assert len(want) >= 7
    for name, h in want.items():
        path = os.path.join(staged, name)
        assert os.path.isfile(path), name
        assert hashlib.sha256(open(path, "rb").read()).hexdigest() == h, f"{name} differs from the committed hash"
        orig = os.path.join("/root/reference", name)
        if os.path.isfile(orig):   # authoring container: also byte-identical to the reference itself
            assert open(orig, "rb").read() == open(path, "rb").read(), name
