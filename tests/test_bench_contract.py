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
    """The CPU arm times the whole pipeline on what it is given and reports those points over that time."""
    b = _bench()
    cloud = b.cpu_sample(60_000, 1, side=2.5)
    assert 0 < len(cloud) < 60_000 and cloud[:, 0].max() < 2.5 and cloud[:, 1].max() < 2.5
    secs, n, n_tiles, n_rows = b.cpu_arm(cloud)
    assert n == len(cloud) and secs > 0 and n_tiles >= 1 and n_rows >= n_tiles * b.CFG["min_pts"]
    base = b.cpu_baseline_dict(n / secs, 8, n, n_tiles, n_rows, secs)
    assert base["kind"] == "port" and base["unit"] == "points/s" and base["cores"] == 8
    assert "nothing extrapolated" in base["sample"] and "KD-tree" in base["sample"]
    from oracle import oracle as O
    assert O.SEARCH == "brute"                                  # the checker's search mode is restored


def test_both_arms_name_the_same_workload_and_defaults_are_small():
    b = _bench()
    assert b.workload(b.N_POINTS) == b.workload(1_000_000)
    assert "1000000-point" in b.workload(b.N_POINTS) and "spatial vote" in b.workload(b.N_POINTS)
    assert b.CFG == dict(grid_size=(2.0, 4.0), min_pts=128, max_pts=16384, batch_size=8, is_wood=0.5)
    assert set(b.load_peaks()) >= {"hbm", "bf16", "bf16_sustained", "src"}
    t = b.load_traffic()
    assert t.get("conv_tc_kernel", 0) > 0 and t.get("grid_query_kernel", 0) > 0
