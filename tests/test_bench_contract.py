"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm's JSON line, the silent non-zero
ranks of that arm under torchrun, the watchdog, and the product arm failing loudly (no CPU fallback) without a device."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(args, env=None, timeout=300):
    e = dict(os.environ)
    e.pop("RANK", None); e.pop("WORLD_SIZE", None); e.pop("LOCAL_RANK", None)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, timeout=timeout, env=e, cwd=ROOT)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "3", "--ref-batch", "4"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_utterances_per_sec" and d["unit"] == "utt/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 3
    assert d["value"] > 0 and abs(d["value"] - 4 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("config2") and d["vs_baseline"] is None and d["data"] == "synthetic"


def test_reference_arm_other_ranks_exit_silently():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_watchdog_ends_a_run_that_does_not_finish():
    r = _run(["--impl", "reference", "--steps", "500", "--warmup", "3", "--watchdog", "2"], timeout=120)
    assert r.returncode != 0
    assert "Timeout" in r.stderr and "most recent call first" in r.stderr      # faulthandler: every thread's Python stack
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_product_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the product arm would simply run")
    r = _run(["--steps", "1", "--warmup", "1", "--no-decode", "--no-cpu-baseline"])
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]      # no line that could be mistaken for a measurement
