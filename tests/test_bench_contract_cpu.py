"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU oracle timed on the
host) must run without CUDA and print ONE JSON line with the driver's keys; the GPU arm must refuse to run
without a device instead of falling back to the oracle."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e,
                          capture_output=True, text=True, timeout=300)


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = _run(["--impl", "reference", "--steps", "2", "--warmup", "1", "--seconds", "20", "--batch", "2", "--cpu-threads", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "input_msamples_per_s_decoded" and d["unit"] == "Msamples/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 2 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]
    assert "configs[3]" in d["config"]["workload"]            # the GPU arm's default workload, same string


def test_reference_arm_single_recording_is_single_threaded_like_the_reference():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--seconds", "20", "--workload", "c2"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == 1 and "configs[1]" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_without_work():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0", "--seconds", "20",
              "--batch", "2"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run(["--steps", "1", "--warmup", "3", "--seconds", "20", "--no-cpu-baseline"])
    assert r.returncode != 0                                  # no CPU fallback for the product path
    assert not any(l.strip().startswith("{") for l in r.stdout.splitlines())
