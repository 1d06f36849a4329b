"""bench.py's CPU-side contract: the reference arm (oracle port of the pipeline) runs on a bounded sample and
reports what the driver expects; both arms describe the same workload.  No GPU needed."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("p2w_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_cpu_arm_on_a_small_plot():
    b = _bench()
    base, est = b.cpu_arm(60_000, 1, budget_s=1.0)
    assert base["kind"] == "port" and base["unit"] == "points/s" and base["cores"] >= 1
    assert base["value"] > 0 and est > 0 and abs(base["value"] - 60_000 / est) < 1e-6 * base["value"]
    assert "batches" in base["sample"] and "spatial vote" in base["sample"]


def test_both_arms_name_the_same_workload_and_defaults_are_small():
    b = _bench()
    assert b.workload(b.N_POINTS) == b.workload(1_000_000)
    assert "1000000-point" in b.workload(b.N_POINTS) and "spatial vote" in b.workload(b.N_POINTS)
    assert b.CFG == dict(grid_size=(2.0, 4.0), min_pts=128, max_pts=16384, batch_size=8, is_wood=0.5)
    assert set(b.load_peaks()) >= {"hbm", "bf16", "bf16_sustained", "src"}
    t = b.load_traffic()
    assert t.get("conv_tc_kernel", 0) > 0 and t.get("grid_query_kernel", 0) > 0
