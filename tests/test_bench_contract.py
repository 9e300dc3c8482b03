"""The JSON lines bench.py printed on the B200 for the final tree (profiles/r2/final_bench_*.json, copied from the GPU box) against the
contract the driver parses: keys, units, the two arms on the same configuration, the roofline / cpu_baseline / e2e objects."""
import json
import os

from conftest import ROOT

R2 = os.path.join(ROOT, "profiles", "r2")


def _load(name):
    return json.load(open(os.path.join(R2, name)))


def test_headline_line_carries_every_contract_key():
    d = _load("final_bench_default_with_planning.json")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["metric"] == "capsule_obstacle_distance_queries_per_s" and d["unit"] == "pair-queries/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["vs_baseline"] is None and d["data"] == "synthetic" and "workload" in d["config"]
    assert d["gpu_launches"] >= d["steps"] > 0
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"]
    r = d["roofline"]
    assert 0 < r["frac"] < 1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0 and r["avg_launch_ms"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == d["unit"] and "sample" in c
    assert d["parity_sample"]["bit_identical"] is True and d["parity_sample"]["verdict_mismatch"] == 0
    assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] > 0.9 * d["clocks"]["sm_max_mhz"]
    rows = d["planning_time_vs_cpu"]
    assert len(rows) == 4 and all("error" not in row for row in rows)                       # BASELINE configs 1-4
    assert all(row["identical_runs_gpu_vs_cpu"] == row["runs"] for row in rows[:3]) and rows[3]["same_run"] is True
    f = d["full_mode"]                                                                      # nearest points of every pair: shipped path vs the one-kernel path, same run
    assert f["split"]["launches"] > f["one_kernel"]["launches"] == 1 and f["split"]["ms_per_launch"] < 0.5 * f["one_kernel"]["ms_per_launch"]


def test_reference_arm_is_the_same_metric_on_the_same_configuration():
    d, r = _load("final_bench_default_with_planning.json"), _load("final_bench_reference.json")
    assert r["impl"] == "reference" and r["gpu_launches"] == 0
    for k in ("metric", "unit", "higher_is_better", "config", "dtype", "scaling"):
        assert d[k] == r[k], k
    assert r["e2e"] == {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert r["cpu_baseline"]["kind"] == "reference" and r["cpu_baseline"]["value"] == r["value"]
    assert d["e2e"]["value"] / r["e2e"]["value"] > 1e4


def test_scaling_lines_are_weak_and_strong_with_a_timed_gather():
    one = _load("r2i_bench_ref32.json")
    for n in (2, 4, 8):
        w, s = _load(f"weak_{n}gpu.json"), _load(f"strong_{n}gpu.json")
        assert w["n_gpus"] == n and w["scaling"] == "weak" and 0.9 * n < w["value"] / one["value"] < 1.05 * n
        assert s["n_gpus"] == n and s["scaling"] == "strong" and s["strong_scaling"]["gather_ms_per_step"] > 0
        assert abs(s["strong_scaling"]["compute_ms_per_step"] + s["strong_scaling"]["gather_ms_per_step"] - s["ms_per_step"]) < 0.2 * s["ms_per_step"]
