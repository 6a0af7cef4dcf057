"""The bench.py JSON contract, checked without a GPU: the reference arm really runs (small workload) and prints the
line the driver parses; the committed B200 line of this round carries every key the contract names."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config"}


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "C1", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"] == "particle_x_point_weight_evals_per_sec" and line["unit"] == "evals/s"
    assert line["value"] > 1e6 and line["higher_is_better"] is True and line["vs_baseline"] is None
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_committed_b200_line_has_every_contract_key():
    line = json.loads((ROOT / "profiles" / "r1_bench_line_n1.json").read_text())
    assert BASE_KEYS <= set(line) and line["n_gpus"] == 1 and line["data"] == "synthetic" and line["dtype"] == "f32"
    assert line["warmup"] >= 3 and line["vs_baseline"] is None and line["scaling"] == "weak"
    rf = line["roofline"]
    assert rf["bound"] in ("hbm", "tensor") and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert rf["traffic"] is None or rf["traffic"] > 0
    cb = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["unit"] == line["unit"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] <= line["value"] * 1.05
    assert line["gpu_launches"] > 0
    ck = line["clocks"]
    assert ck["sm_mhz"] and ck["sm_max_mhz"] and not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert "workload" in line["config"] and "l2" in line["config"]
