"""CPU: the pieces of bench.py that do not need a GPU -- the roofline arithmetic, the workload description and the
reference arm's JSON line (contract keys), on a tiny sample so the whole file runs in seconds."""
import json
import subprocess
import sys
import types
from pathlib import Path

import bench

ROOT = Path(__file__).resolve().parents[1]


def test_roofline_entry_int8_path_uses_flop_weighted_product_count():
    eng = types.SimpleNamespace(fp64_slices=7, fp64_slices_kinv=6, fp64_min_dim=512)
    r = bench.roofline_entry(eng, 75.0, 37.0, {"bf16_tflops_sustained": 1408.0}, "MEASURED_PEAKS.json", 1e12, 100, 1000.0, 1)
    pairs_eff = (28 + 2 * 21) / 3.0
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s"
    assert abs(r["peak"] - 2 * 1408.0 / pairs_eff) < 1e-9
    assert abs(r["frac"] - 75.0 / r["peak"]) < 1e-12
    assert r["traffic"] and r["vs_fp64_dmma_peak"] == 75.0 / 37.0


def test_roofline_entry_pure_fp64_mode_is_measured_against_the_dmma_peak():
    eng = types.SimpleNamespace(fp64_slices=0, fp64_slices_kinv=6, fp64_min_dim=512)
    r = bench.roofline_entry(eng, 32.0, 37.0, {}, "fallback", 1e12, 100, 1000.0, 1)
    assert r["peak"] == 37.0 and abs(r["frac"] - 32.0 / 37.0) < 1e-12


def test_workload_config_names_the_baseline_configuration():
    args = types.SimpleNamespace(n=0, workload="c2", gpus=1)
    cfg = bench.workload_config(args, bench.WORKLOADS["c2"], 1)
    assert cfg["n"] == 44484 and cfg["d"] == 21 and cfg["latents_per_gpu"] == 4 and "model" not in cfg
    assert "AdamW" in cfg["step"]


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--points", "300"], capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
