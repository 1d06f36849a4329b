"""The C-ABI library loads without a GPU and exports every symbol include/p2w.h declares; the
Python binding table mirrors the header one to one.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "p2w.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(p2w_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def handle():
    from pointstowood_b200 import build
    build.build()
    return ctypes.CDLL(os.path.join(ROOT, "pointstowood_b200", "libp2w.so"))


def test_header_declares_the_hot_path_entry_points():
    names = _declared()
    for must in ("p2w_knn", "p2w_radius", "p2w_fps", "p2w_grid", "p2w_pointnet_conv_max", "p2w_scatter_minmax",
                 "p2w_pack", "p2w_writeback", "p2w_sort_pairs", "p2w_unique_last"):
        assert must in names


def test_library_exports_every_declared_symbol(handle):
    missing = [n for n in _declared() if not hasattr(handle, n)]
    assert not missing, f"declared in include/p2w.h but not exported: {missing}"


def test_binding_table_matches_header():
    from pointstowood_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "p2w.h")).read(), flags=re.S)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\((.*?)\)\s*;", header, flags=re.S)
        assert m, name
        params = [a for a in m.group(1).split(",") if a.strip() and a.strip() != "void"]
        assert len(params) == len(argtypes), f"{name}: header has {len(params)} parameters, binding {len(argtypes)}"


def test_version_and_error_slot_work_without_gpu(handle):
    handle.p2w_version.restype = ctypes.c_int
    handle.p2w_last_error.restype = ctypes.c_char_p
    assert handle.p2w_version() >= 100
    assert isinstance(handle.p2w_last_error(), bytes)


def test_argument_errors_are_reported_before_any_launch(handle):
    """Bad arguments return P2W_EINVAL with a message and never reach the device."""
    handle.p2w_knn.restype = ctypes.c_int
    handle.p2w_last_error.restype = ctypes.c_char_p
    rc = handle.p2w_knn(None, None, None, None, ctypes.c_int32(1), ctypes.c_int64(0), ctypes.c_int64(4),
                        ctypes.c_int32(1000), None, None, None)
    assert rc == -1 and b"k=1000" in handle.p2w_last_error()
