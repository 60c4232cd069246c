"""The JSON lines bench.py owes the driver (CPU-checkable parts).

The reference arm (`--impl reference`) needs no GPU: it times the oracle port of the reference's CPU path.  The native arm's
line is checked on the committed run of the same script (profiles/), which the GPU box produced."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "cpu_baseline"}


def _check_common(d):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["unit"] == "graphs/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                      # BASELINE.md holds no published number for this metric
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["warmup"] >= 3 and d["value"] > 0 and d["ms_per_step"] > 0
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"]) and d["cpu_baseline"]["kind"] in ("port", "reference")


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    _check_common(d)
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1
    assert d["dtype"] == "f64"                           # the reference computes in fp64 (gnnLightning.py:L939)


def test_committed_native_line_has_the_contract_shape():
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_tc_v7_bench.json")))
    _check_common(d)
    assert d["config"]["workload"] == "mini_cheetah-k4-contact" and d["config"]["global_batch"] == 16384 * d["n_gpus"]
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"]                # host buffers + PCIe inside the timed region
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"])
    ro = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(ro) and ro["bound"] in ("hbm", "tensor")
    assert abs(ro["frac"] - ro["achieved"] / ro["peak"]) < 1e-9
    # the per-kernel shares are measured live in the same run (CUDA events around every launch) against the profiled step
    # time, which also holds the gaps between launches: they account for (nearly) the whole step
    assert 0.9 < sum(k["share"] for k in d["kernels"]) <= 1.0 + 1e-9
