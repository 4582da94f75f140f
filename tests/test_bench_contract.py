"""bench.py prints exactly one JSON line on stdout with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
             "data", "config", "e2e", "cpu_baseline"}


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["unit"] == "Msamples/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_our_arm_line_small_batch():
    d = _run(["--steps", "2", "--warmup", "1"], env={"PMR446_BENCH_STREAMS": "16"})
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline", "kernels", "chain_roofline"} <= set(d)
    assert d["n_gpus"] == 1 and d["dtype"] == "f32" and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert d["gpu_launches"] >= 2 * 5 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 16 * 2400000 * 2 and d["e2e"]["d2h_bytes_per_step"] == 16 * 16 * 12500 * 2
    r = d["roofline"]
    assert r["bound"] in ("hbm", "fp32") and {"achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0 < r["frac"] < 5
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
