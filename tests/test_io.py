"""pointstowood_b200.io against files written by the reference's own src/io.py (oracle/make_golden_io.py) and
against itself: byte-identical PLY / PCD writers, readers for binary (both byte orders) and ascii bodies,
chunked writing, error behaviour.  CPU only."""
import os

import numpy as np
import pandas as pd
import pytest

from pointstowood_b200 import io as pio


def _frame(n=257, seed=11):                # the frame oracle/make_golden_io.py wrote
    rng = np.random.default_rng(seed)
    df = pd.DataFrame({"x": rng.normal(size=n).astype(np.float32), "y": rng.normal(size=n), "z": rng.normal(size=n),
                       "reflectance": rng.normal(-8, 2, size=n).astype(np.float32),
                       "red": rng.integers(0, 255, n), "green": rng.integers(0, 255, n), "blue": rng.integers(0, 255, n),
                       "label": rng.integers(0, 2, n).astype(np.float64), "pwood": rng.random(n)})
    df["note"] = "leaf"
    return df


def test_write_ply_is_byte_identical_to_the_reference(golden_dir, tmp_path):
    out = tmp_path / "a.ply"
    pio.write_ply(str(out), _frame(), comments=["p2w golden"])
    assert out.read_bytes() == open(os.path.join(golden_dir, "io_ref.ply"), "rb").read()


def test_write_ply_in_chunks_gives_the_same_bytes(golden_dir, tmp_path, monkeypatch):
    monkeypatch.setattr(pio, "CHUNK_ROWS", 100)          # 257 rows: two full chunks and a ragged one
    out = tmp_path / "b.ply"
    pio.write_ply(str(out), _frame(), comments=["p2w golden"])
    assert out.read_bytes() == open(os.path.join(golden_dir, "io_ref.ply"), "rb").read()


def test_read_ply_of_the_reference_file(golden_dir):
    df, want = pio.read_ply(os.path.join(golden_dir, "io_ref.ply")), _frame()
    assert list(df.columns) == ["x", "y", "z", "red", "green", "blue", "reflectance", "label", "pwood"]
    for c in df.columns:
        assert np.array_equal(df[c].to_numpy(), want[c].to_numpy().astype(df[c].dtype)), c
    assert df["red"].dtype == np.int32 and df["x"].dtype == np.float64
    pc, extra = pio.load_file(os.path.join(golden_dir, "io_ref.ply"), additional_headers=True)
    assert extra == ["red", "green", "blue", "reflectance", "label", "pwood"] and len(pc) == 257


def test_pcd_writer_and_reader(golden_dir, tmp_path):
    out = tmp_path / "a.pcd"
    pio.write_pcd(_frame().rename(columns={"reflectance": "intensity"}), str(out))
    assert out.read_bytes() == open(os.path.join(golden_dir, "io_ref.pcd"), "rb").read()
    df = pio.read_pcd(str(out))
    assert list(df.columns) == ["x", "y", "z", "intensity"] and len(df) == 257
    assert np.array_equal(df["intensity"].to_numpy(), _frame()["reflectance"].to_numpy())


def test_ascii_and_big_endian_ply(tmp_path):
    rows = np.array([[0.5, 1.5, -2.0, 7.0], [1.0, 2.0, 3.0, -4.0]])
    a = tmp_path / "a.ply"
    a.write_text("ply\nformat ascii 1.0\nelement vertex 2\nproperty float x\nproperty float y\nproperty float z\n"
                 "property float reflectance\nend_header\n" + "\n".join(" ".join(str(v) for v in r) for r in rows) + "\n")
    assert np.array_equal(pio.read_ply(str(a)).to_numpy(), rows)
    b = tmp_path / "b.ply"
    with open(b, "wb") as f:
        f.write(b"ply\nformat binary_big_endian 1.0\nelement vertex 2\nproperty double x\nproperty double y\n"
                b"property double z\nproperty uchar reflectance\nend_header\n")
        rec = np.zeros(2, dtype=[("x", ">f8"), ("y", ">f8"), ("z", ">f8"), ("reflectance", "u1")])
        for i, name in enumerate(rec.dtype.names):
            rec[name] = np.abs(rows[:, i])
        rec.tofile(f)
    got = pio.read_ply(str(b))
    assert np.array_equal(got.to_numpy(dtype=np.float64), np.abs(rows)) and got["reflectance"].dtype == np.uint8


def test_save_file_round_trip_and_errors(tmp_path):
    arr = np.random.default_rng(0).normal(size=(50, 5))
    out = tmp_path / "plot_ours.ply"
    pio.save_file(str(out), arr, additional_fields=["label", "pwood"])
    back = pio.load_file(str(out))
    assert list(back.columns) == ["x", "y", "z", "label", "pwood"] and np.array_equal(back.to_numpy(), arr)
    mesh = tmp_path / "mesh.ply"
    mesh.write_text("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nelement face 1\n"
                    "property list uchar int vertex_indices\nend_header\n0\n3 0 0 0\n")
    with pytest.raises(Exception, match="mesh"):
        pio.read_ply(str(mesh))
    with pytest.raises(Exception, match="not recognised"):
        pio.load_file(str(tmp_path / "cloud.xyz"))


def test_predict_column_handling():
    from pointstowood_b200.predict import preprocess_point_cloud_data
    df = pd.DataFrame({"X": [0.0], "Y": [1.0], "Z": [2.0], "Red": [1], "scalar_Intensity": [3.0], "label": [1.0]})
    out, headers, has = preprocess_point_cloud_data(df)
    assert list(out.columns) == ["x", "y", "z", "reflectance", "red"] and headers == ["red", "reflectance"] and has
    out, headers, _ = preprocess_point_cloud_data(pd.DataFrame({"x": [0.0], "y": [1.0], "z": [2.0]}))
    assert list(out.columns) == ["x", "y", "z", "reflectance"] and out["reflectance"][0] == 0.0 and headers == []


def test_predict_flags_match_the_reference(golden_dir):
    """Every flag of the reference's predict.py exists with the same spellings, default, type, nargs and action
    (tests/golden/predict_flags.json is read from the reference's source by oracle/make_golden_cli.py)."""
    import json
    from pointstowood_b200.predict import build_parser
    want = json.load(open(os.path.join(golden_dir, "predict_flags.json")))
    have = {}
    for a in build_parser()._actions:
        if a.option_strings:
            have[max(a.option_strings, key=len)] = a
    for flag, spec in want.items():
        assert flag in have, f"missing {flag}"
        a = have[flag]
        assert sorted(a.option_strings) == spec["names"]
        if "default" in spec:
            assert a.default == spec["default"], flag
        if "type" in spec:
            assert a.type.__name__ == spec["type"], flag
        if "nargs" in spec:
            assert a.nargs == spec["nargs"], flag
        if spec.get("action") == "store_true":
            assert a.const is True and a.nargs == 0, flag
