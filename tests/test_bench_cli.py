"""bench.py on a machine without a GPU: the reference arm must run (it only needs the host cores and oracle/_ref or the
oracle port) and print the contract's JSON line; the product arm must fail loudly — there is no CPU fallback to time."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    out = _run("--impl", "reference", "--workload", "config2", "--steps", "1", "--warmup", "0")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Mtriangles/s" and d["unit"] == "Mtri/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["config"]["gridsize"] == 1024 and d["config"]["triangles"] == 5110 and d["config"]["mode"] == "surface"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mtri/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_has_no_cpu_fallback():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the product arm runs")
    out = _run("--steps", "1", "--warmup", "1", "--workload", "config2", "--no-cpu-baseline")
    assert out.returncode != 0
    assert not any(l.strip().startswith("{") for l in out.stdout.splitlines()), "no bench line may be printed without a GPU"
