"""CPU: the bench.py contract.  The reference arm (the compiled reference on the host cores) prints
ONE JSON line with the agreed keys; the product arm refuses to run without a GPU (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


def test_reference_arm_prints_one_contract_line(ref):
    p = _run("--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-elems", "20000")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "samples/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("PG(h,z) samples/sec") and d["config"]["workload"].startswith("cfg4")
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f64" and d["data"] == "synthetic"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["value"] == d["value"] and e2e["unit"] == d["unit"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0


@pytest.mark.parametrize("config,unit,elems", [("cfg1", "samples/s", "1000000"), ("cfg2", "samples/s", "20000"),
                                               ("cfg3-gamma", "samples/s", "2000"), ("cfg3-saddle", "samples/s", "20000"),
                                               ("cfg5a-pdf", "points/s", "10000"), ("cfg5a-logcdf", "points/s", "2500")])
def test_reference_arm_covers_every_config(ref, config, unit, elems):
    """bench.py --config <BASELINE config> --impl reference: the reference's own call for that config."""
    p = _run("--impl", "reference", "--config", config, "--steps", "1", "--warmup", "1", "--ref-elems", elems)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == unit and d["value"] > 0
    assert d["config"]["workload"].startswith(config.split("-")[0])
    assert d["cpu_baseline"]["kind"] == "reference" and d["e2e"]["value"] == d["value"]


def test_product_arm_needs_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = _run("--steps", "1", "--warmup", "1", "--no-cpu-baseline", "--no-e2e", timeout=300)
    assert p.returncode != 0
    assert not any(ln.lstrip().startswith("{") for ln in p.stdout.splitlines())  # no bench line from a CPU
