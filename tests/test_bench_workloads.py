"""bench.py's four workloads at toy sizes on the GPU: the line's contract keys, the in-run sanity checks (histogram
totals == pixels, the exchange merged every frame of the step), the reduced-mosaic parity check of c4 and the
sustained / e2e / ceiling legs -- so that a change to the engine cannot silently break a workload the driver does
not run by default."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _bench(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    return json.loads(lines[0])


@pytest.mark.timeout(900)
@pytest.mark.parametrize("workload,extra", [
    ("c2", ["--frames", "3", "--sustain-s", "0.05"]),
    ("c3", ["--total-frames", "20", "--group", "8"]),                    # two full groups and a remainder of 4
    ("c4", ["--total-frames", "6"]),
    ("c5", ["--total-frames", "700", "--group", "256", "--ring-groups", "2"]),   # the ring wraps, remainder of 188
])
def test_bench_workload_line(workload, extra):
    d = _bench("--workload", workload, "--steps", "2", "--warmup", "3", "--height", "96", "--width", "128",
               "--cpu-frames", "1", *extra)
    assert d["metric"].startswith("RGNir Mpix/s") and d["unit"] == "Mpix/s" and d["n_gpus"] == 1 and d["steps"] == 2
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["scaling"] == ("weak" if workload == "c2" else "strong")
    assert d["dtype"] == ("u16" if workload == "c3" else "u8")
    assert d["config"]["height"] == 96 and d["config"]["width"] == 128 and "workload" in d["config"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert r["algorithmic_bytes_per_px"] == (30 if workload == "c3" else 27)
    assert d["gpu_launches"] > 0 and "reasons" in d["clocks"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["ranks"]["exchange_tail_us"][0] >= 0
    if workload == "c2":
        assert d["sustained"]["steps"] >= 2 and d["sustained"]["value"] > 0 and "single_frame" in d
    if workload == "c4":
        assert d["parity_check"]["ok"] is True and d["collectives"]["wb_hist_all_reduce_us"]["n"] == 2
    if workload in ("c3", "c5"):
        assert d["config"]["launch_groups_per_step_per_gpu"] == 3
