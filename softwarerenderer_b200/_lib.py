"""Loader of the C-ABI library (libswr_b200.so).  There is no CPU fallback: if the CUDA extension
has not been built the import of any product module fails loudly."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# SWR_LIB_VARIANT selects an experimental build (make VARIANT=_x EXTRA=-D...); unset = the product
LIB_PATH = os.path.join(HERE, "libswr_b200" + os.environ.get("SWR_LIB_VARIANT", "") + ".so")

MAX_RENDER_TARGETS = 12
MAX_UNIFORM_BYTES = 1024
MAX_MIP_LEVELS = 14


class SwrStats(C.Structure):
    _fields_ = [
        ("fragments", C.c_uint64), ("primitives_in", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("draws", C.c_uint64), ("passes", C.c_uint64),
        ("last_geometry_ms", C.c_float), ("last_tile_ms", C.c_float),
        ("last_tile_size", C.c_int32), ("reserved", C.c_int32), ("scratch_bytes", C.c_uint64), ("h2d_bytes", C.c_uint64),
    ]


class StockUniforms(C.Structure):
    """struct swr_stock_uniforms (include/swr_b200.h)."""
    _fields_ = [("mvp", C.c_float * 16), ("texture", C.c_void_p), ("tex_w", C.c_int32), ("tex_h", C.c_int32),
                ("mip", C.c_void_p * MAX_MIP_LEVELS), ("mip_w", C.c_int32 * MAX_MIP_LEVELS), ("mip_h", C.c_int32 * MAX_MIP_LEVELS),
                ("mip_levels", C.c_int32), ("max_anisotropy", C.c_int32)]


# every symbol include/swr_b200.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("swr_stock_vertex_shader", _P, [C.c_int]),
    ("swr_stock_pixel_shader", _P, [C.c_int]),
    ("swr_create", C.c_int, [C.POINTER(_P), C.c_int]),
    ("swr_destroy", None, [_P]),
    ("swr_last_error", C.c_char_p, []),
    ("swr_abi_version", C.c_int, []),
    ("swr_set_viewport", C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("swr_set_depth_range", C.c_int, [_P, C.c_float, C.c_float]),
    ("swr_set_cull_mode", C.c_int, [_P, C.c_int]),
    ("swr_set_vertex_attrib_pointer", C.c_int, [_P, C.c_int, C.c_int, _P, C.c_size_t]),
    ("swr_set_vertex_shader", C.c_int, [_P, _P]),
    ("swr_set_raster_mode", C.c_int, [_P, C.c_int]),
    ("swr_set_scissor_rect", C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("swr_set_pixel_shader", C.c_int, [_P, _P]),
    ("swr_set_render_target", C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int]),
    ("swr_set_uniforms", C.c_int, [_P, _P, C.c_size_t]),
    ("swr_set_tile_size", C.c_int, [_P, C.c_int]),
    ("swr_set_tile_split", C.c_int, [_P, C.c_int]),
    ("swr_set_index_narrowing", C.c_int, [_P, C.c_int]),
    ("swr_debug_pack_indices16", C.c_int, [_P, C.c_size_t, _P, _P]),
    ("swr_set_tile_partition", C.c_int, [_P, C.c_int, C.c_int]),
    ("swr_set_scratch_limit", C.c_int, [_P, C.c_size_t]),
    ("swr_set_stream", C.c_int, [_P, _P]),
    ("swr_set_pipeline", C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    ("swr_draw_elements", C.c_int, [_P, C.c_int, C.c_size_t, _P]),
    ("swr_draw_raster_list", C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, C.c_size_t]),
    ("swr_finish", C.c_int, [_P]),
    ("swr_process_elements", C.c_int, [_P, C.c_int, C.c_size_t, _P, _P, _P]),
    ("swr_get_stats", C.c_int, [_P, C.POINTER(SwrStats)]),
    ("swr_reset_stats", C.c_int, [_P]),
    ("swr_timer_begin", C.c_int, [_P]),
    ("swr_timer_end", C.c_int, [_P, C.POINTER(C.c_float)]),
    ("swr_device_alloc", _P, [_P, C.c_size_t]),
    ("swr_device_free", C.c_int, [_P, _P]),
    ("swr_host_alloc_pinned", _P, [C.c_size_t]),
    ("swr_host_free_pinned", C.c_int, [_P]),
    ("swr_memcpy_h2d", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("swr_memcpy_d2h", C.c_int, [_P, _P, _P, C.c_size_t]),
    ("swr_memset32", C.c_int, [_P, _P, C.c_uint32, C.c_size_t]),
    ("swr_flush_l2", C.c_int, [_P]),
    ("swr_pack_tiles", C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    ("swr_unpack_tiles", C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    ("swr_owned_tile_count", C.c_int64, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("swr_set_tile_mirrors", C.c_int, [_P, C.c_int, C.c_int, _P]),
    ("swr_ipc_get_handle", C.c_int, [_P, _P, C.POINTER(C.c_int64)]),
    ("swr_ipc_open", C.c_int, [_P, _P, C.c_int64, C.POINTER(C.c_void_p)]),
    ("swr_ipc_close", C.c_int, [_P, _P]),
    ("swr_debug_enable_tile_stats", C.c_int, [_P, C.c_int]),
    ("swr_debug_read_tile_stats", C.c_int64, [_P, _P, C.c_int64]),
    ("swr_shared_scratch_create", C.c_int, [_P, C.c_size_t, C.POINTER(C.c_void_p)]),
    ("swr_set_geometry_shards", C.c_int, [_P, C.c_int, C.c_int, _P]),
    ("swr_peer_barrier", C.c_int, [_P]),
]

STREAM_OUT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t)

_lib = None


class SwrError(RuntimeError):
    pass


def load():
    """dlopen libswr_b200.so and bind every declared symbol.  Raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()' "
                f"or make -C softwarerenderer_b200/csrc).  softwarerenderer_b200 has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(lib, name)      # AttributeError if the library does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise SwrError(f"{what or 'swr call'} failed ({rc}): {load().swr_last_error().decode()}")
