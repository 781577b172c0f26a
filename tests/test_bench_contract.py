"""The JSON line `bench.py` prints is a contract with the round-end driver: both arms are run here
at a tiny size and the keys / types the driver reads are checked."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric": str, "value": (int, float), "unit": str, "n_gpus": int, "steps": int, "warmup": int,
          "ms_per_step": (int, float), "higher_is_better": bool, "scaling": str, "dtype": str, "data": str, "config": dict,
          "e2e": dict, "gpu_launches": int}


def run_bench(*flags):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, "exactly ONE JSON line on stdout"
    return json.loads(lines[0])


def check_common(d):
    for k, t in COMMON.items():
        assert k in d and isinstance(d[k], t), (k, d.get(k))
    assert "vs_baseline" in d and d["vs_baseline"] is None        # BASELINE.md holds no published number for this metric
    assert d["metric"] == "svmpc_control_steps_per_sec" and d["scaling"] == "weak" and d["higher_is_better"] is True
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k


def test_reference_arm_line():
    """`--impl reference`: the CPU port of the reference's path, no GPU needed."""
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "3")
    check_common(d)
    assert d["impl"] == "reference" and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0


@pytest.mark.gpu
def test_device_arm_line():
    d = run_bench("--steps", "3", "--warmup", "3", "--instances", "256", "--cpu-sample", "2")
    check_common(d)
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] >= 3 and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["gpu_launches"] == d["steps"], "one fused kernel per control step in the timed region"
    assert d["e2e"]["value"] > 0 and d["e2e"]["value"] != d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and "sample" in cb
    c = d["clocks"]
    assert "sm_mhz" in c and "sm_max_mhz" in c and isinstance(c["reasons"], list)
    assert "l2_policy" in d["config"]
    # the other half of BASELINE's metric rides in the same line: large-N phi with its own roofline, clocks and in-run checks
    ph = d["phi"]
    assert ph["N"] == 65536 and ph["d"] == 40 and ph["ms_phi"] > 0 and ph["ms_phi_with_median"] > ph["ms_phi"]
    pr = ph["roofline"]
    assert pr["kernel"] == "phi_tc_kernel" and pr["bound"] == "tensor" and abs(pr["frac"] - pr["achieved"] / pr["peak"]) < 1e-9
    assert pr["frac_vs_inrun_cublas_tf32"] > 0 and "sm_mhz" in ph["clocks"]
    assert ph["rel_err_vs_float64_rows"] <= 1e-4 and ph["median"]["ulp_distance"] <= 4
    cf = d["configs"]
    for name in ("pendulum_demo", "particle_demo", "dual_stress"):
        assert cf[name]["device_ms_per_dual_step"] > 0 and cf[name]["wall_ms_per_dual_step"] > 0, name
    for name in ("pendulum_demo", "particle_demo"):
        assert cf[name]["drop_in_classes"]["wall_ms_per_dual_step"] > 0
        assert cb["configs"][name]["ms_per_dual_step"] > 0
