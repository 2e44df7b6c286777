"""bench.py prints ONE JSON line with the keys the driver reads (task statement, section 4).  The reference arm runs on
the CPU; our arm needs a B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "e2e", "cpu_baseline"}


def run_bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert COMMON <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "frames_per_sec_720x576" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["dtype"] == "u8" and d["vs_baseline"] is None and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


@pytest.mark.gpu
def test_our_arm_line():
    d = run_bench("--steps", "2", "--warmup", "3")
    assert COMMON | {"gpu_launches", "roofline", "clocks", "kernel_ms_per_step"} <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "weak"
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["value"] != d["value"]
    # continuous streams: only the 64 NEW frames of each of the 16 streams travel per step
    assert d["e2e"]["h2d_bytes_per_step"] == 16 * 64 * 720 * 576 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0 and d["gpu_launches"] % (2 * 4) == 0          # 2 steps x 4 launch sequences of the same kernels
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] > 0
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # the same workload description in both arms
    ref = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0")
    assert ref["config"] == d["config"] and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
