"""bench.py's output contract: exactly ONE JSON line on stdout carrying the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(*flags, timeout=600):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_prints_one_json_line():
    d = _run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "2", "--blocks", "1", "--sims", "20")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "mcts_sims_per_sec" and d["unit"] == "sims/s"
    from oracle import build_ref
    want = "reference" if build_ref.available() else "port"      # the unmodified reference (oracle/_ref) whenever it is staged
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == want and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert ("UNMODIFIED reference" in d["cpu_baseline"]["sample"]) == (want == "reference")
    assert d["e2e"] == {"value": d["value"], "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["higher_is_better"] is True


@pytest.mark.gpu
def test_gpu_arm_prints_one_json_line():
    d = _run("--steps", "1", "--warmup", "3", "--ticks-per-step", "24", "--fused-chunk", "12", "--cpu-seconds", "2", "--no-aux-rooflines")
    assert BASE_KEYS | {"gpu_launches", "clocks", "roofline"} <= set(d)
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["dtype"] == "f16" and d["scaling"] == "weak"
    assert d["fused_tick_kernel"] is True and d["gpu_launches"] == 2 and d["e2e"]["value"] > 0     # 24 ticks = 2 fused launches of 12 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    assert r["bound"] == "tensor" and 0 < r["frac"] < 1.2 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["unit"] == "TFLOP/s"
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["value"] > 0 and c["cores"] >= 1 and c["sample"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
