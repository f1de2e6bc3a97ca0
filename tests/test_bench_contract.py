"""bench.py's one-JSON-line contract: the reference arm on the CPU (runs everywhere), our arm on the GPU (`-m gpu`)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"}


def _run(args, timeout, env=None):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout,
                         cwd=ROOT, env=dict(os.environ, **(env or {})))
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line_uses_every_core_even_under_torchrun_env():
    """`--impl reference`: oracle port on the host cores, same metric / config as our arm; torchrun's OMP_NUM_THREADS=1 must
    not turn it into a single-thread baseline (the N >= 2 reference ratios were void for that reason)."""
    line = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], 600, env={"OMP_NUM_THREADS": "1"})
    assert BASE_KEYS <= set(line) and line["impl"] == "reference"
    assert line["metric"].startswith("graph-pairs/sec") and line["unit"] == "graph-pairs/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("eval_batch synthetic sequence of 1000 graphs")


def test_reference_arm_is_silent_on_other_ranks():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.gpu
def test_our_arm_line_carries_every_key():
    line = _run(["--steps", "20", "--warmup", "3", "--no-cpu-baseline"], 600)
    assert BASE_KEYS <= set(line) and "impl" not in line
    assert line["n_gpus"] == 1 and line["gpu_launches"] == 20 and line["dtype"] == "f32"
    assert line["config"]["batch"] == 128 and line["config"]["node_num"] == 64 and line["config"]["k"] == 20
    roof = line["roofline"]
    assert roof["bound"] == "hbm" and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and roof["traffic"]
    assert abs(roof["achieved"] - 128 * 8196 / (line["ms_per_step"] * 1e-3) / 1e9) / roof["achieved"] < 1e-6
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 2 * 128 * 15 * 64 * 4 and e2e["d2h_bytes_per_step"] == 512 and 0 < e2e["value"] < line["value"]
    assert {"c_abi_host_call", "pageable_inputs", "compact_inputs"} <= set(e2e)
    assert line["clocks"]["sm_mhz"] and not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    scan = line["scan"]
    assert scan["graphs"] == 4000 and scan["max_abs_diff_vs_fused_pair_kernel"] <= 5e-6
    assert set(scan["phases_ms"]) >= {"embed_row_block", "allgather_pooled", "score_row_block_tcgen05"}
    assert line["train"]["launches_per_step"] == 10 and line["train"]["ms_per_step"] > 0
    assert {(r["node_num"], r["k"]) for r in line["sweep"]["rows"]} == {(16, 10), (32, 10), (32, 20), (64, 10), (64, 20), (128, 10), (128, 20)}
    small = line["small_batch"]
    assert 0 < small["B1_kernel_us"] <= small["B37_kernel_us"] * 1.1 and small["B37_kernel_us"] < line["ms_per_step"] * 1e3
    assert small["one_pair_e2e_us"] > small["B1_kernel_us"]


def test_profiles_index_names_existing_files():
    """profiles/current.json (read by bench.py for the roofline's measured DRAM traffic) names captures that are committed."""
    with open(os.path.join(ROOT, "profiles", "current.json")) as f:
        index = json.load(f)
    names = [n for v in index.values() for n in (v if isinstance(v, list) else [v])]
    missing = [n for n in names if not os.path.exists(os.path.join(ROOT, "profiles", n))]
    assert not missing, missing
    with open(os.path.join(ROOT, "profiles", index["embed_ncu_summary"])) as f:
        assert json.load(f)["dram_bytes_per_launch"] > 0
