"""ctypes binding of libp2w.so (the C ABI declared in include/p2w.h).

There is deliberately NO fallback: if the library is missing or a call fails, the op
raises.  The CPU oracle under oracle/ is test infrastructure and is never imported here.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libp2w.so")

_P = c_void_p
# name -> (restype, argtypes); mirrors include/p2w.h one to one
SIGNATURES = {
    "p2w_version": (c_int32, []),
    "p2w_last_error": (c_char_p, []),
    "p2w_device_info": (c_int32, [_P, _P, _P]),
    "p2w_launch_count": (ctypes.c_longlong, []),
    "p2w_knn": (c_int32, [_P, _P, _P, _P, c_int32, c_int64, c_int64, c_int32, _P, _P, _P]),
    "p2w_radius": (c_int32, [_P, _P, _P, _P, c_int32, c_int64, c_int64, c_double, c_int32, _P, _P, _P]),
    "p2w_grid_search_ws_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "p2w_grid_search_pair_evals": (c_int32, [_P, c_int64, c_int64, c_int32, _P]),
    "p2w_knn_grid": (c_int32, [_P, _P, _P, _P, c_int32, c_int64, c_int64, c_int32, _P, _P, _P, c_size_t, _P]),
    "p2w_knn_grid_ex": (c_int32, [_P, _P, _P, _P, c_int32, c_int64, c_int64, c_int32, c_float, c_int32, _P, _P, _P, c_size_t, _P]),
    "p2w_spatial_vote": (c_int32, [_P, c_int64, c_int32, _P, _P, c_float, _P, _P, _P]),
    "p2w_radius_grid": (c_int32, [_P, _P, _P, _P, c_int32, c_int64, c_int64, c_double, c_int32, _P, _P, _P, c_size_t,
                                  _P]),
    "p2w_table_count": (c_int32, [_P, c_int64, c_int32, _P, _P]),
    "p2w_table_to_edges": (c_int32, [_P, c_int64, c_int32, _P, c_int64, _P, _P]),
    "p2w_fps": (c_int32, [_P, _P, _P, c_int32, c_int64, _P, _P, _P]),
    "p2w_colminmax": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P, _P]),
    "p2w_grid": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P, _P, _P, _P, _P]),
    "p2w_voxel_keys_grouped": (c_int32, [_P, c_int64, c_int32, _P, c_int32, _P, c_int32, c_float, c_int32, _P, _P, _P,
                                         _P, _P]),
    "p2w_sort_ws_bytes": (c_size_t, [c_int64]),
    "p2w_sort_pairs": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, _P, _P]),
    "p2w_unique_ws_bytes": (c_size_t, [c_int64]),
    "p2w_unique_last": (c_int32, [_P, _P, c_int64, _P, _P, _P, _P, _P, _P]),
    "p2w_pointnet_conv_max": (c_int32, [_P, _P, _P, _P, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32,
                                        _P, _P, _P, _P, _P, _P, _P, c_int32, _P, c_size_t, _P]),
    "p2w_pointnet_conv_max_ex": (c_int32, [_P, c_int32, _P, _P, _P, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32,
                                           _P, _P, _P, _P, _P, _P, _P, c_int32, c_int32, _P, c_size_t, c_int32, _P, _P]),
    "p2w_pointnet_conv_ws_bytes": (c_size_t, [c_int32, c_int32, c_int32, c_int32]),
    "p2w_knn_interpolate": (c_int32, [_P, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, _P]),
    "p2w_knn_interpolate_ex": (c_int32, [_P, c_int32, _P, _P, _P, c_int64, c_int32, c_int32, c_int32, _P, c_int32, _P]),
    "p2w_knn_interpolate_cat": (c_int32, [_P, c_int32, _P, _P, _P, c_int64, c_int32, c_int32, _P, c_int32, c_int32, c_int32,
                                          _P, c_int32, _P]),
    "p2w_knn_interpolate_add": (c_int32, [_P, c_int32, _P, _P, _P, c_int64, c_int32, c_int32, _P, _P, c_int32, c_int32, _P]),
    "p2w_affine_relu": (c_int32, [_P, _P, c_int64, c_int32, _P, _P, _P, _P, c_int32, _P]),
    "p2w_dense_expand_ws_bytes": (c_size_t, [c_int32, c_int32]),
    "p2w_dense_expand": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P, _P, _P, _P, _P, c_size_t, c_int32, _P]),
    "p2w_rowdot": (c_int32, [_P, c_int32, c_int64, c_int32, _P, c_float, _P, _P]),
    "p2w_add_relu": (c_int32, [_P, _P, _P, c_int64, c_int32, _P]),
    "p2w_segment_max": (c_int32, [_P, _P, c_int32, c_int32, _P, _P]),
    "p2w_segment_max_ex": (c_int32, [_P, c_int32, _P, c_int32, c_int32, _P, _P, _P, _P]),
    "p2w_scatter_minmax": (c_int32, [_P, _P, c_int64, c_int32, c_int64, c_int32, _P, _P, _P]),
    "p2w_sa_prepare": (c_int32, [_P, c_int32, _P, _P, _P, c_int32, c_int64, _P, _P, _P]),
    "p2w_pack": (c_int32, [_P, c_int32, _P, _P, c_int32, c_int64, _P, _P, _P, _P, _P, _P]),
    "p2w_ground_normalize": (c_int32, [_P, c_int32, c_int64, _P, c_float, c_int32, c_int32, _P, _P, _P]),
    "p2w_reflectance_keys": (c_int32, [_P, c_int32, c_int32, c_int64, _P, _P]),
    "p2w_reflectance_normalize": (c_int32, [_P, c_int64, _P, _P, _P, _P]),
    "p2w_reflectance_values": (c_int32, [_P, c_int64, c_int64, c_int64, _P, _P, _P]),
    "p2w_reflectance_scale": (c_int32, [_P, c_int64, _P, _P, _P]),
    "p2w_assemble5": (c_int32, [_P, c_int32, _P, _P, c_int64, _P, _P]),
    "p2w_sampling_keys": (c_int32, [_P, c_int32, c_int32, _P, _P, c_int64, c_float, ctypes.c_uint32, _P, _P]),
    "p2w_replacement_picks": (c_int32, [_P, _P, _P, c_int32, c_int32, ctypes.c_uint64, _P, _P]),
    "p2w_ground_min": (c_int32, [_P, c_int32, c_int64, _P, c_float, c_int32, c_int32, _P, _P]),
    "p2w_ground_apply": (c_int32, [_P, c_int32, c_int64, _P, c_float, c_int32, c_int32, _P, _P, _P]),
    "p2w_writeback": (c_int32, [_P, _P, _P, _P, c_int32, c_int64, c_float, _P, _P, _P, _P, _P]),
}

_lib = None


class P2WError(RuntimeError):
    """Raised for the cases upstream reports through TORCH_CHECK / AT_ASSERTM."""


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -m pointstowood_b200.build` "
                "(there is no CPU or PyTorch fallback for these ops)")
        handle = ctypes.CDLL(SO_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def to_device(values, device, dtype=None):
    """Small host metadata (CSR pointers, plans) to the device WITHOUT blocking the host: staged through the
    caching pinned allocator and copied asynchronously.  torch.tensor(list, device=...) copies from pageable
    memory, which makes the host wait for everything queued on the stream before it can launch again."""
    import numpy as np
    import torch
    a = np.ascontiguousarray(values if dtype is None else np.asarray(values, dtype=dtype))
    return torch.from_numpy(a).pin_memory().to(device, non_blocking=True)


def check(rc: int) -> None:
    if rc != 0:
        raise P2WError(lib().p2w_last_error().decode() or f"libp2w error {rc}")
