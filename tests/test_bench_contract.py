"""The bench line the driver reads: every key of the contract is present in the line recorded at the end of the round
(profiles/bench_r02zd_default.json, written by `python bench.py` on a B200), and its numbers are consistent with each other."""
import glob
import json
import os

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _line():
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_r02z*_default.json")))
    assert paths, "no recorded bench line under profiles/"
    with open(paths[-1]) as f:
        return json.loads(f.read().strip().splitlines()[-1])


def test_recorded_bench_line_has_the_contract_keys():
    d = _line()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert "workload" in d["config"] and "model" not in d["config"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in d["roofline"], k
    for k in ("value", "unit", "cores", "kind", "sample"):
        assert k in d["cpu_baseline"], k
    for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert k in d["e2e"], k
    for k in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert k in d["clocks"], k


def test_recorded_bench_line_is_self_consistent():
    d = _line()
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    # achieved = algorithmic bytes per launch / average launch duration
    alg = r["algorithmic_bytes_per_cell_update"] * r["cells_per_launch"]
    assert abs(r["achieved"] - alg / (r["avg_launch_ms"] * 1e-3) / 1e9) < 1e-6 * r["achieved"]
    # value = cell-updates of K steps / time of K steps
    nzp, nxp = d["config"]["padded_grid"]
    updates = 2.0 * nzp * nxp * d["config"]["nt"] * d["config"]["shots_total"]
    assert abs(d["value"] - updates / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * d["value"]
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] != d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
