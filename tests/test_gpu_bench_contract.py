"""GPU: bench.py's own arm prints one JSON line carrying every key of the measurement contract (short run)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_b200_arm_json_line():
    env = dict(os.environ, BF_C5_MINUTES="0.5", BF_C5_STREAM_MINUTES="2")      # the C5 leg streams 2 min instead of 1 h
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "3", "--warmup", "3", "--frames",
                          "16", "--no-cpu"], capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["value"] > 0 and line["gpu_launches"] >= 3
    rf = line["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in rf, key
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert "workload" in line["config"] and "model" not in line["config"]
    for extra in ("miso", "mvdr", "replay", "heatmap", "fir"):
        assert line[extra] and "error" not in line[extra], (extra, line[extra])
    stream = line["replay"]["one_hour_stream"]
    assert stream["frames"] == 4 * 900 and stream["h2d_gb"] > 5.0 and stream["pcie_h2d_gb_per_s"] > 1.0
    assert set(stream["stage_ms"]) == {"h2d", "ingest", "maps", "overlay"}
    assert line["mvdr"]["roofline"]["frac"] > 0.2
