"""Golden files for tests/test_io.py, written by the REFERENCE's own src/io.py (run in the build container, where
/root/reference exists):  python oracle/make_golden_io.py
    tests/golden/io_ref.ply   write_ply of a seeded frame (x y z reflectance red green blue label pwood + a text column)
    tests/golden/io_ref.pcd   write_pcd of the same frame (x y z intensity)
TEST INFRASTRUCTURE: nothing in the product imports this."""
import os
import sys

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference/pointstowood")
from src import io as ref_io  # noqa: E402


def frame(n=257, seed=11):
    rng = np.random.default_rng(seed)
    df = pd.DataFrame({"x": rng.normal(size=n).astype(np.float32), "y": rng.normal(size=n), "z": rng.normal(size=n),
                       "reflectance": rng.normal(-8, 2, size=n).astype(np.float32),
                       "red": rng.integers(0, 255, n), "green": rng.integers(0, 255, n), "blue": rng.integers(0, 255, n),
                       "label": rng.integers(0, 2, n).astype(np.float64), "pwood": rng.random(n)})
    df["note"] = "leaf"                      # not convertible to float: the writer must skip it silently
    return df


if __name__ == "__main__":
    out = os.path.join(ROOT, "tests", "golden")
    df = frame()
    ref_io.write_ply(os.path.join(out, "io_ref.ply"), df.copy(), comments=["p2w golden"])
    ref_io.write_pcd(df.rename(columns={"reflectance": "intensity"}).copy(), os.path.join(out, "io_ref.pcd"))
    back = ref_io.read_ply(os.path.join(out, "io_ref.ply"))
    print(back.dtypes.to_dict(), len(back))
