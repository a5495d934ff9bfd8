"""ctypes binding of include/uwcuda.h.  Loads the in-tree libuwcuda.so and fails loudly if it
is missing -- there is no CPU fallback and nothing here touches oracle/."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .build import LIB_PATH

UW_OK = 0
UW_ERR_INVALID, UW_ERR_CUDA, UW_ERR_NO_DEVICE, UW_ERR_OOM, UW_ERR_NOT_READY, UW_ERR_UNSUPPORTED = 1, 2, 3, 4, 5, 6

FLAG_EXACT_F64 = 0x1
FLAG_INDEX32 = 0x2
FLAG_KEEP_DENSITIES = 0x4
FLAG_TRIS = 0x8
FLAG_STAGED = 0x10
FLAG_ORDERED = 0x20
FLAG_ANALYTIC_SKIP = 0x40
FLAG_EXPORTABLE = 0x80

CHUNK_BLANK_EARLY = 0x1
CHUNK_HAS_MESH = 0x2
CHUNK_U16_OVERFLOW = 0x4


class UwConfig(C.Structure):
    _fields_ = [
        ("internal_size", C.c_int32), ("chunk_size", C.c_int32), ("octaves", C.c_uint32),
        ("iso_level", C.c_float), ("max_height", C.c_float), ("adj_z_mod", C.c_float),
        ("min_hue", C.c_float), ("max_hue", C.c_float), ("saturation", C.c_float), ("base_value", C.c_float),
        ("min_z", C.c_float), ("max_z", C.c_float),
        ("seed", C.c_uint32), ("device", C.c_int32), ("flags", C.c_uint32), ("guard_eps", C.c_float),
        ("reserved", C.c_uint32 * 4),
    ]


class UwBatchView(C.Structure):
    _fields_ = [
        ("n_chunks", C.c_uint32), ("n_verts", C.c_uint64), ("n_inds", C.c_uint64),
        ("descs", C.c_void_p), ("verts", C.c_void_p), ("inds16", C.c_void_p), ("inds32", C.c_void_p),
        ("tris", C.c_void_p), ("tri_cell_start", C.c_void_p),
    ]


class UwDeviceView(C.Structure):
    _fields_ = [
        ("n_chunks", C.c_uint32), ("n_verts", C.c_uint64), ("n_inds", C.c_uint64),
        ("d_descs", C.c_void_p), ("d_verts", C.c_void_p), ("d_inds16", C.c_void_p), ("d_inds32", C.c_void_p),
        ("d_densities", C.c_void_p), ("density_stride", C.c_uint32),
    ]


class UwStageTimes(C.Structure):
    _fields_ = [
        ("noise_ms", C.c_float), ("classify_ms", C.c_float), ("scan_ms", C.c_float), ("emit_ms", C.c_float),
        ("total_ms", C.c_float), ("launches", C.c_uint32),
    ]


UW_MAX_SEGMENTS = 16
GATHER_DESCS_TO_HOST = 0x1
GATHER_DRAW_TO_HOST = 0x2


class UwGatherInfo(C.Structure):
    """uw_gather_info: plain bytes, handed from the rendering process to the producers."""
    _fields_ = [
        ("abi_version", C.c_uint32), ("n_segments", C.c_uint32), ("device", C.c_int32), ("index_bytes", C.c_uint32),
        ("owner_pid", C.c_uint64), ("base", C.c_uint64), ("bytes", C.c_uint64),
        ("off_head", C.c_uint64), ("off_descs", C.c_uint64), ("off_verts", C.c_uint64), ("off_inds", C.c_uint64), ("off_draw", C.c_uint64),
        ("n_chunks", C.c_uint64), ("seg_vcap", C.c_uint64), ("seg_icap", C.c_uint64),
        ("ipc_handle", C.c_uint8 * 64),
    ]


class UwGatherSegment(C.Structure):
    _fields_ = [
        ("first_chunk", C.c_uint64), ("n_chunks", C.c_uint32), ("n_mesh", C.c_uint32), ("n_blank", C.c_uint32),
        ("overflow", C.c_uint32), ("n_verts", C.c_uint64), ("n_inds", C.c_uint64), ("guard", C.c_uint64),
    ]


class UwGatherResult(C.Structure):
    _fields_ = [
        ("n_segments", C.c_uint32), ("epoch", C.c_uint32),
        ("n_chunks", C.c_uint64), ("n_verts", C.c_uint64), ("n_inds", C.c_uint64),
        ("d_descs", C.c_void_p), ("d_verts", C.c_void_p), ("d_inds", C.c_void_p),
        ("seg_vcap", C.c_uint64), ("seg_icap", C.c_uint64),
        ("h_descs", C.c_void_p), ("h_draw", C.c_void_p), ("d_draw", C.c_void_p), ("n_draw", C.c_uint64),
        ("seg", UwGatherSegment * UW_MAX_SEGMENTS),
    ]


DESC_DTYPE = np.dtype([("pos", "<i4", (3,)), ("flags", "<u4"), ("vert_offset", "<u4"), ("vert_count", "<u4"),
                       ("index_offset", "<u4"), ("index_count", "<u4")])
VERT_DTYPE = np.dtype([("pos", "<f4", (3,)), ("color", "<f4", (3,))])
TRI_DTYPE = np.dtype([("verts", "<f4", (3, 3)), ("normal", "<f4", (3,))])
assert DESC_DTYPE.itemsize == 32 and VERT_DTYPE.itemsize == 24 and TRI_DTYPE.itemsize == 48

# every symbol include/uwcuda.h declares (tests check the library exports all of them)
EXPORTS = [
    "uw_abi_version", "uw_config_default", "uw_create", "uw_destroy", "uw_last_error", "uw_perm_table",
    "uw_build", "uw_build_async", "uw_batch_wait", "uw_batch_view_get", "uw_batch_free",
    "uw_build_device", "uw_sync", "uw_device_view_get",
    "uw_debug_densities", "uw_debug_cases", "uw_build_from_densities", "uw_iso_at",
    "uw_set_stream", "uw_get_stage_times", "uw_set_profiling", "uw_get_guard_count", "uw_debug_ffma_peak", "uw_export_arena_fd", "uw_debug_vertex_colors",
    "uw_gather_create", "uw_gather_destroy", "uw_gather_attach", "uw_gather_detach", "uw_gather_build",
    "uw_gather_build_device", "uw_gather_wait", "uw_slab_bounds",
    "uw_multi_create", "uw_multi_build", "uw_multi_destroy", "uw_multi_last_error", "uw_debug_copy_to_host", "uw_raycast_tris", "uw_slab_bounds_weighted", "uw_multi_render_share", "uw_share_search_next",
]

_lib = None


class UwShareSearch(C.Structure):
    """uw_share_search (include/uwcuda.h): state of the render-share search; zero-initialised = not started."""
    _fields_ = [("share", C.c_double), ("step", C.c_double), ("last_cost", C.c_double), ("best_share", C.c_double),
                ("best_cost", C.c_double), ("dir", C.c_int32), ("moves", C.c_uint32), ("settled", C.c_uint32), ("reserved", C.c_uint32)]


class UwError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"uwcuda status {status}: {msg}")
        self.status = status


def load_library() -> C.CDLL:
    """Load libuwcuda.so from the package tree.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("UWCUDA_LIB", LIB_PATH)      # override only for A/B experiments of kernel variants
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -m underwaterworld_b200.build` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = C.CDLL(path)
    vp, u32, i32p = C.c_void_p, C.c_uint32, C.c_void_p
    lib.uw_abi_version.restype = C.c_uint32
    lib.uw_config_default.argtypes = [C.POINTER(UwConfig)]
    lib.uw_config_default.restype = None
    lib.uw_create.argtypes = [C.POINTER(UwConfig), C.POINTER(vp)]
    lib.uw_destroy.argtypes = [vp]
    lib.uw_destroy.restype = None
    lib.uw_last_error.argtypes = [vp]
    lib.uw_last_error.restype = C.c_char_p
    lib.uw_perm_table.argtypes = [vp, vp]
    lib.uw_build.argtypes = [vp, i32p, u32, C.POINTER(vp)]
    lib.uw_build_async.argtypes = [vp, i32p, u32, C.POINTER(vp)]
    lib.uw_batch_wait.argtypes = [vp]
    lib.uw_batch_view_get.argtypes = [vp, C.POINTER(UwBatchView)]
    lib.uw_batch_free.argtypes = [vp]
    lib.uw_batch_free.restype = None
    lib.uw_build_device.argtypes = [vp, vp, u32]
    lib.uw_sync.argtypes = [vp]
    lib.uw_device_view_get.argtypes = [vp, C.POINTER(UwDeviceView)]
    lib.uw_debug_densities.argtypes = [vp, i32p, u32, vp]
    lib.uw_debug_cases.argtypes = [vp, i32p, u32, vp]
    lib.uw_build_from_densities.argtypes = [vp, i32p, vp, u32, C.POINTER(vp)]
    lib.uw_iso_at.argtypes = [vp, vp, u32, vp]
    lib.uw_set_stream.argtypes = [vp, vp]
    lib.uw_get_stage_times.argtypes = [vp, C.POINTER(UwStageTimes)]
    lib.uw_set_profiling.argtypes = [vp, C.c_int]
    lib.uw_get_guard_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    lib.uw_debug_ffma_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.uw_debug_vertex_colors.argtypes = [vp, vp, vp, u32, vp]
    lib.uw_export_arena_fd.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_uint64)]
    u64 = C.c_uint64
    lib.uw_gather_create.argtypes = [vp, u32, u64, u64, u64, C.POINTER(UwGatherInfo)]
    lib.uw_gather_destroy.argtypes = [vp]
    lib.uw_gather_attach.argtypes = [vp, C.POINTER(UwGatherInfo), u32]
    lib.uw_gather_detach.argtypes = [vp]
    lib.uw_gather_build.argtypes = [vp, i32p, u32, u64]
    lib.uw_gather_build_device.argtypes = [vp, vp, u32, u64]
    lib.uw_gather_wait.argtypes = [vp, u32, C.POINTER(UwGatherResult)]
    lib.uw_slab_bounds.argtypes = [u32, u32, u32, C.POINTER(u32), C.POINTER(u32)]
    lib.uw_slab_bounds.restype = None
    lib.uw_slab_bounds_weighted.argtypes = [u32, u32, u32, u32, u32, C.POINTER(u32), C.POINTER(u32)]
    lib.uw_slab_bounds_weighted.restype = None
    lib.uw_multi_render_share.argtypes = [vp]
    lib.uw_multi_render_share.restype = C.c_uint32
    lib.uw_share_search_next.argtypes = [C.POINTER(UwShareSearch), u32, C.c_double]
    lib.uw_share_search_next.restype = C.c_double
    lib.uw_multi_create.argtypes = [C.POINTER(UwConfig), C.POINTER(C.c_int32), u32, C.POINTER(vp)]
    lib.uw_multi_build.argtypes = [vp, i32p, u32, u32, C.POINTER(UwGatherResult)]
    lib.uw_multi_destroy.argtypes = [vp]
    lib.uw_multi_destroy.restype = None
    lib.uw_multi_last_error.argtypes = [vp]
    lib.uw_multi_last_error.restype = C.c_char_p
    lib.uw_debug_copy_to_host.argtypes = [vp, u64, vp]
    lib.uw_raycast_tris.argtypes = [vp, vp, vp, u32, C.c_int32, vp]
    _lib = lib
    return lib
