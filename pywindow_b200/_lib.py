"""ctypes binding of ``csrc/libpwb.so`` (C ABI declared in ``include/pwb.h``).

There is no CPU fallback: if the shared library is missing or cannot be loaded the import of any
compute entry point raises.  Build it with ``python __graft_entry__.py`` (or ``make -C
pywindow_b200/csrc``).
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PWB_LIBRARY: an alternative build of the same sources (kernel-variant experiments, tools/variants.sh)
LIB_PATH = os.environ.get("PWB_LIBRARY") or os.path.join(_HERE, "csrc", "libpwb.so")

MAX_TYPES = 16
REC_STATUS, REC_NATOMS, REC_COM, REC_MAXD, REC_MAXD_A1, REC_MAXD_A2, REC_AVGD = 0, 1, 2, 5, 6, 7, 8
REC_PORE_D, REC_PORE_ATOM, REC_OPT_D, REC_OPT_ATOM, REC_OPT_C = 9, 10, 11, 12, 13
REC_NRAYS, REC_NCLEAR, REC_NWIN, REC_EPS, REC_OPT_NIT, REC_OPT_STOP, REC_RADIUS, REC_NSURV = 16, 17, 18, 19, 20, 21, 22, 23
REC_OPT_NFEV = 24
REC_WIN0 = 28
ST_OPT_BOUNDS, ST_NO_CLEAR_RAY, ST_RAY_CAP, ST_WIN_CAP, ST_WIN_DROPPED, ST_WIN_NEGATIVE, ST_BAD_ITEM = 1, 2, 4, 8, 16, 32, 64
RAY_TILE = 1024
STAGE_SCALARS, STAGE_AVGD, STAGE_OPT, STAGE_RAYS, STAGE_CLUSTER, STAGE_WINDOWS, STAGE_ALL = 1, 2, 4, 8, 16, 32, 63
OPTION_OPT_NO_LIST = 1          # PWB_OPTION_OPT_NO_LIST

EXPORTS = [
    "pwb_abi_version", "pwb_last_error", "pwb_ctx_create", "pwb_ctx_destroy", "pwb_ctx_set_chunks", "pwb_ctx_set_option", "pwb_workspace_bytes",
    "pwb_workspace_layout", "pwb_ray_capacity", "pwb_frame_scalars", "pwb_average_diameter",
    "pwb_opt_pore", "pwb_ray_search", "pwb_cluster", "pwb_window_refine", "pwb_full_analysis_batch",
    "pwb_full_analysis_host", "pwb_full_analysis_host_async", "pwb_launch_count", "pwb_fma_peak", "pwb_dm_workspace_bytes",
    "pwb_dm_workspace_layout", "pwb_discrete_molecules", "pwb_dm_compact", "pwb_set_error", "pwb_history_index",
    "pwb_history_decode", "pwb_xyz_index", "pwb_xyz_decode", "pwb_pdb_decode", "pwb_history_write", "pwb_stream_probe",
]


class Batch(ctypes.Structure):
    _fields_ = [
        ("n_items", ctypes.c_int32), ("max_atoms", ctypes.c_int32), ("n_types", ctypes.c_int32),
        ("shared_types", ctypes.c_int32), ("ray_cap", ctypes.c_int32), ("win_cap", ctypes.c_int32),
        ("rec_stride", ctypes.c_int32), ("stages", ctypes.c_int32),
        ("adjust", ctypes.c_double), ("increment", ctypes.c_double), ("increment2", ctypes.c_double),
        ("pore_opt", ctypes.c_int32), ("total_atoms", ctypes.c_int32),
        ("coords", ctypes.c_void_p), ("item_off", ctypes.c_void_p), ("atom_type", ctypes.c_void_p),
        ("type_vdw", ctypes.c_void_p), ("type_mass", ctypes.c_void_p), ("centre_override", ctypes.c_void_p),
        ("opt_start", ctypes.c_void_p), ("opt_bounds", ctypes.c_void_p),
        ("records", ctypes.c_void_p), ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_size_t),
        ("ray_tile_begin", ctypes.c_int32), ("ray_tile_end", ctypes.c_int32),
    ]


class WsLayout(ctypes.Structure):
    _fields_ = [(n, ctypes.c_size_t) for n in
                ("atoms", "atom_idx", "item_d", "item_i", "clear_k", "clear_pos", "clear_a2", "labels",
                 "win_k", "big_pts", "big_i", "atom_scratch", "opt_hist", "total")]


class PwbLibraryError(RuntimeError):
    pass


_lib = None


def load():
    """Load libpwb.so (once) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PwbLibraryError(
            f"{LIB_PATH} not built -- run `python __graft_entry__.py` (nvcc, sm_100a). "
            "pywindow_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    L.pwb_abi_version.restype = ci
    L.pwb_last_error.restype = ctypes.c_char_p
    L.pwb_ctx_create.argtypes = [ctypes.POINTER(vp), ci, ci]
    L.pwb_ctx_destroy.argtypes = [vp]
    L.pwb_ctx_set_chunks.argtypes = [vp, ci]
    L.pwb_ctx_set_option.argtypes = [vp, ci, ci]
    L.pwb_ctx_set_option.restype = ci
    L.pwb_workspace_bytes.argtypes = [ci, ci, ci, ci, ci]
    L.pwb_workspace_bytes.restype = ctypes.c_size_t
    L.pwb_workspace_layout.argtypes = [ci, ci, ci, ci, ci, ctypes.POINTER(WsLayout)]
    L.pwb_ray_capacity.argtypes = [vp, ctypes.POINTER(Batch), vp, ctypes.POINTER(ci)]
    for name in ("pwb_frame_scalars", "pwb_average_diameter", "pwb_opt_pore", "pwb_ray_search",
                 "pwb_cluster", "pwb_window_refine", "pwb_full_analysis_batch"):
        getattr(L, name).argtypes = [vp, ctypes.POINTER(Batch), vp]
        getattr(L, name).restype = ci
    L.pwb_full_analysis_host.argtypes = [vp, ctypes.POINTER(Batch), vp, vp, vp, vp, ci, vp]
    L.pwb_full_analysis_host.restype = ci
    L.pwb_full_analysis_host_async.argtypes = [vp, ctypes.POINTER(Batch), vp, vp, vp, vp, ci, vp]
    L.pwb_full_analysis_host_async.restype = ci
    L.pwb_dm_workspace_bytes.argtypes = [ci, ci, ci]
    L.pwb_dm_workspace_bytes.restype = ctypes.c_size_t
    L.pwb_discrete_molecules.argtypes = [ci, ci, ci, ctypes.c_double, ctypes.c_double, vp, vp, ci, vp, ci, vp, vp,
                                         vp, ci, ci, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    L.pwb_discrete_molecules.restype = ci
    L.pwb_dm_compact.argtypes = [ci, ci, ci, ci] + [vp] * 14
    L.pwb_dm_compact.restype = ci
    i64 = ctypes.c_int64
    L.pwb_history_index.argtypes = [vp, i64, vp, i64]
    L.pwb_history_index.restype = i64
    L.pwb_history_decode.argtypes = [vp, i64, vp, vp, i64, ci, ci, ci, vp, vp, ci]
    L.pwb_history_decode.restype = i64
    L.pwb_xyz_index.argtypes = [vp, i64, vp, vp, vp, i64]
    L.pwb_xyz_index.restype = i64
    for name in ("pwb_xyz_decode", "pwb_pdb_decode"):
        getattr(L, name).argtypes = [vp, i64, vp, vp, i64, ci, vp, ci]
        getattr(L, name).restype = i64
    L.pwb_stream_probe.argtypes = [vp, i64, ci]
    L.pwb_stream_probe.restype = i64
    L.pwb_history_write.argtypes = [ctypes.c_char_p, ctypes.c_char_p, vp, vp, vp, i64, ci, vp, ci, ci, ci]
    L.pwb_history_write.restype = i64
    L.pwb_fma_peak.argtypes = [vp, ci, vp, ctypes.POINTER(ctypes.c_double)]
    L.pwb_fma_peak.restype = ci
    L.pwb_launch_count.argtypes = [vp]
    L.pwb_launch_count.restype = ctypes.c_longlong
    if L.pwb_abi_version() != 4:
        raise PwbLibraryError("libpwb.so ABI version mismatch")
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise PwbLibraryError(load().pwb_last_error().decode())
