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
    line = json.loads((ROOT / "profiles" / "r2_bench_line_n1.json").read_text())
    assert BASE_KEYS <= set(line) and line["n_gpus"] == 1 and line["data"] == "synthetic" and line["dtype"] == "f32"
    assert line["warmup"] >= 3 and line["vs_baseline"] is None and line["scaling"] == "weak"
    # one full roofline block per particle distribution of SURVEY 8(d): T (tracking cluster) is bound by the L1/L2 gather
    # rate, G (global) by DRAM lines; both ceilings are measured in the same run and the fraction follows from the line
    for key, bound in (("roofline", "l2-gather"), ("roofline_G", "hbm-line")):
        rf = line[key]
        assert rf["bound"] == bound and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
        assert abs(rf["achieved"] - rf["algorithmic_bytes_per_launch"] / (rf["update_ms"] * 1e-3) / 1e9) < 1e-6 * rf["achieved"]
        assert rf["traffic"] is None or rf["traffic"] > 0
        assert rf["traffic_key"].startswith("C3 ")
    peaks = line["gather_peaks_loads_per_sec"]
    assert abs(line["roofline"]["peak"] - peaks["whole_grid_coherent"] * 32 / 1e9) < 1e-6 * line["roofline"]["peak"]
    assert abs(line["roofline_G"]["peak"] - line["roofline_G"]["hbm_peak"] / 4) < 1e-6 * line["roofline_G"]["peak"]
    cb = line["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(cb) and cb["unit"] == line["unit"]
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0 and 0 < e2e["value"] <= line["value"] * 1.05
    cpp = line["e2e_cpp_class"]
    assert cpp["set_uploads_per_step"] == 1.0 and cpp["set_downloads_per_step"] == 1.0 and 0 < cpp["value"] <= line["value"]
    assert line["gpu_launches"] > 0
    ck = line["clocks"]
    assert ck["sm_mhz"] and ck["sm_max_mhz"] and not set(ck["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert "workload" in line["config"] and "l2" in line["config"]
    # BASELINE configs 1 and 2 and the config-4 strong-scaling point ride in the same driver-run line
    assert line["configs"]["C1"]["cpu_reference_1core"]["cores"] == 1 and line["configs"]["C1"]["ms_per_frame"] > 0
    c2 = line["configs"]["C2"]
    assert c2["roofline"]["bound"] == "hbm" and 0 < c2["roofline"]["frac"] < 1 and c2["voxels_per_sec"] > 1e10
    assert line["secondary_c4_strong"]["scaling"] == "strong"


def test_reference_and_b200_lines_name_the_same_config():
    a = json.loads((ROOT / "profiles" / "r2_bench_line_n1.json").read_text())
    b = json.loads((ROOT / "profiles" / "r2_bench_line_ref.json").read_text())
    assert a["config"] == b["config"] and a["metric"] == b["metric"] and a["unit"] == b["unit"]
    assert b["impl"] == "reference" and a["steps"] == b["steps"] and a["warmup"] == b["warmup"]


def test_committed_multi_gpu_lines_assert_parity():
    seen = 0
    for n in (2, 4, 8):
        path = ROOT / "profiles" / f"r2_bench_line_n{n}.json"
        if not path.exists():
            continue
        seen += 1
        line = json.loads(path.read_text())
        assert line["n_gpus"] == n and line["parity_n"] == "bit-identical" and line["scaling"] == "weak"
        assert line["config"]["particles_total"] == n * 100_000
        assert line["secondary_c4_strong"]["n_gpus"] == n
        assert 0 <= line["secondary_fast_resample"]["indices_differing_from_exact"] <= line["secondary_fast_resample"]["of"]
    assert seen >= 2
