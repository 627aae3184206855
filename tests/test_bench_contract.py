"""bench.py's driver contract, checked without a GPU: the reference arm runs here (it is the reference's
CPU renderer) and must print exactly one JSON line with the contract's keys; the newest committed line of
the CUDA arm (profiles/*_bench.json, written on a B200) must carry them too."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_reference_arm_prints_one_contract_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "ref_render")):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], check=True, capture_output=True, text=True, cwd=ROOT)
    lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, res.stdout
    line = json.loads(lines[0])
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["impl"] == "reference" and line["metric"] == "Mrays/s" and line["unit"] == "Mrays/s"
    assert line["config"]["workload"].startswith("cfg1") and line["dtype"] == "f64" and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and line["ms_per_step"] > 0


def test_newest_committed_cuda_line_has_the_contract_keys():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench.json")))
    assert files
    line = json.load(open(files[-1]))
    assert BASE_KEYS | {"roofline", "clocks"} <= set(line), (BASE_KEYS | {"roofline", "clocks"}) - set(line)
    assert line["metric"] == "Mrays/s" and line["n_gpus"] == 1 and line["higher_is_better"] is True
    assert line["scaling"] == "weak" and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert line["gpu_launches"] == line["steps"]  # one render kernel per step in the timed region
    roof = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof)
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    e2e = line["e2e"]
    assert e2e["d2h_bytes_per_step"] == 1920 * 1080 * 3 and e2e["h2d_bytes_per_step"] > 0 and e2e["frame_ok"]
    assert min(e2e["passes_s"]) >= 0.3 and e2e["d2h_only"]["GBps"] > 0  # a pass is long enough to be repeatable
    assert roof["frac"] <= 1.0  # executed flops over the measured peak: a bound by construction
    cpu = line["cpu_baseline"]
    assert cpu["kind"] in ("reference", "port") and cpu["cores"] >= 1 and cpu["value"] > 0 and cpu["sample"]
    clocks = line["clocks"]
    assert clocks["sm_mhz"] > 0 and not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(clocks["reasons"])
    # value and ms_per_step describe the same thing
    assert abs(line["value"] - line["workload_stats"]["rays_per_step"] / (line["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * line["value"]
    # both arms describe the job with the same `config`
    ref = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_bench_reference.json")))
    assert json.load(open(ref[-1]))["config"] == line["config"]
