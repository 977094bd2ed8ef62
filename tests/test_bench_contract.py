"""The reference arm of bench.py runs without a GPU: check the JSON line contract here (the GPU arm's line has the
same keys plus roofline / clocks and is produced on the B200 box; profiles/r2_bench_n1.json holds the last one)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, GSG_CPU_BUDGET_S="1.0")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "d4k3n7",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "DOF-updates/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "sample" in line["cpu_baseline"] and line["config"]["workload"]


def test_committed_gpu_bench_line_has_contract_keys():
    with open(os.path.join(ROOT, "profiles", "r2_bench_n1.json")) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["dtype"] == "f64" and line["n_gpus"] == 1 and line["gpu_launches"] > 0
    r = line["roofline"]
    assert r["bound"] == "hbm" and 0 < r["frac"] < 1.5 and r["unit"] == "GB/s"
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    sm = r["step_model"]
    assert sm["bytes_per_dof_model"] == 528 and sm["taylor"]["bytes_per_dof"] == 432 and sm["staged"]["ms_per_step"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and "COMPLETE" in line["cpu_baseline"]["sample"]


def test_committed_multi_gpu_and_recon_lines():
    for n in (2, 4, 8):
        with open(os.path.join(ROOT, "profiles", f"r2_bench_n{n}.json")) as f:
            line = json.loads(f.read().strip().splitlines()[-1])
        assert line["n_gpus"] == n and line["scaling"] == "strong" and line["gpu_launches"] > 0
        assert "gsg_mg" in line["config"]["parallelism"] and line["e2e"]["value"] > 0
    with open(os.path.join(ROOT, "profiles", "r2_bench_recon.json")) as f:
        line = json.loads(f.read().strip().splitlines()[-1])
    assert line["unit"] == "points/s" and line["roofline"]["bound"] == "fp64" and line["cpu_baseline"]["max_abs_diff_vs_gpu"] < 1e-12
