"""ctypes front end of the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  Nothing under softwarerenderer_b200/ does.

  run(scene, impl="oracle")  -> our C restatement   (oracle/libswr_oracle.so)
  run(scene, impl="ref")     -> the unmodified reference build (oracle/_ref/libswr_ref.so)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libswr_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libswr_ref.so")
VARY_PLANES = 8
STREAM_FLOATS = 16


class SwrScene(C.Structure):
    """Mirror of struct swr_scene (oracle/swr_scene.h)."""
    _fields_ = [
        ("vertices", C.c_void_p), ("stride", C.c_int32), ("num_vertices", C.c_int32),
        ("indices", C.c_void_p), ("index_count", C.c_int64),
        ("draw_mode", C.c_int32), ("cull_mode", C.c_int32), ("raster_mode", C.c_int32),
        ("vp_x", C.c_int32), ("vp_y", C.c_int32), ("vp_w", C.c_int32), ("vp_h", C.c_int32),
        ("sc_x", C.c_int32), ("sc_y", C.c_int32), ("sc_w", C.c_int32), ("sc_h", C.c_int32),
        ("depth_n", C.c_float), ("depth_f", C.c_float),
        ("vs_kind", C.c_int32), ("ps_kind", C.c_int32),
        ("mvp", C.c_float * 16),
        ("texture", C.c_void_p), ("tex_w", C.c_int32), ("tex_h", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32),
        ("color", C.c_void_p), ("depth", C.c_void_p), ("count", C.c_void_p),
        ("prim_id", C.c_void_p), ("vary", C.c_void_p),
        ("stream", C.c_void_p), ("stream_cap", C.c_int64), ("stream_len", C.c_int64),
        ("fragments", C.c_uint64), ("primitives_out", C.c_uint64),
    ]


def build(force: bool = False) -> None:
    """make -C oracle (C restatement always; reference build when /root/reference exists)."""
    if force or not os.path.exists(ORACLE_SO) or (os.path.isdir("/root/reference") and not os.path.exists(REF_SO)) \
            or _stale():
        subprocess.run(["make", "-s", "-C", HERE, "all"], check=True)


def _stale() -> bool:
    try:
        so = os.path.getmtime(ORACLE_SO)
        return any(os.path.getmtime(os.path.join(HERE, f)) > so for f in ("swr_oracle.c", "swr_scene.h"))
    except OSError:
        return True


_libs = {}


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def lib(impl: str):
    if impl not in _libs:
        path = ORACLE_SO if impl == "oracle" else REF_SO
        l = C.CDLL(path)
        prefix = "oracle" if impl == "oracle" else "ref"
        fn = getattr(l, prefix + "_draw")
        fn.argtypes = [C.POINTER(SwrScene)]
        fn.restype = C.c_int
        fr = getattr(l, prefix + "_draw_raster_triangles")
        fr.argtypes = [C.POINTER(SwrScene), C.c_void_p, C.c_int64]
        fr.restype = C.c_int
        fp = getattr(l, prefix + "_draw_raster_prims")
        fp.argtypes = [C.POINTER(SwrScene), C.c_int, C.c_void_p, C.c_int64]
        fp.restype = C.c_int
        _libs[impl] = l
    return _libs[impl]


def fresh_targets(width: int, height: int):
    """Cleared render targets (same clear values as softwarerenderer_b200.api.RenderTargets)."""
    n = width * height
    return {
        "color": np.zeros(n, dtype=np.uint32),
        "depth": np.ones(n, dtype=np.float32),
        "count": np.zeros(n, dtype=np.uint32),
        "prim_id": np.full(n, 0xFFFFFFFF, dtype=np.uint32),
        "vary": np.zeros(VARY_PLANES * n, dtype=np.float32),
    }


def _fill(scene, targets, stream_cap: int):
    s = SwrScene()
    keep = [scene.vertices, scene.indices]
    s.vertices = scene.vertices.ctypes.data
    s.stride = scene.stride
    s.num_vertices = scene.num_vertices
    s.indices = scene.indices.ctypes.data
    s.index_count = int(scene.indices.size)
    s.draw_mode, s.cull_mode, s.raster_mode = scene.draw_mode, scene.cull_mode, scene.raster_mode
    s.vp_x, s.vp_y, s.vp_w, s.vp_h = scene.viewport
    s.sc_x, s.sc_y, s.sc_w, s.sc_h = scene.scissor
    s.depth_n, s.depth_f = scene.depth_range
    s.vs_kind, s.ps_kind = scene.vs, scene.ps
    s.mvp = (C.c_float * 16)(*[float(x) for x in scene.mvp.reshape(-1)])
    if scene.texture is not None:
        tex = np.ascontiguousarray(scene.texture, dtype=np.uint32)
        keep.append(tex)
        s.texture = tex.ctypes.data
        s.tex_h, s.tex_w = tex.shape
    s.width, s.height = scene.width, scene.height
    for k in ("color", "depth", "count", "prim_id", "vary"):
        setattr(s, k, targets[k].ctypes.data)
    stream = None
    if stream_cap:
        stream = np.zeros((stream_cap, STREAM_FLOATS), dtype=np.float32)
        s.stream = stream.ctypes.data
        s.stream_cap = stream_cap
    return s, keep, stream


def run(scene, impl: str = "oracle", targets: Optional[dict] = None, stream_cap: int = 0) -> dict:
    """Draw `scene` on the CPU; returns the targets plus fragments / primitives_out / seconds."""
    targets = fresh_targets(scene.width, scene.height) if targets is None else targets
    s, keep, stream = _fill(scene, targets, stream_cap)
    fn = getattr(lib(impl), ("oracle" if impl == "oracle" else "ref") + "_draw")
    t0 = time.perf_counter()
    rc = fn(C.byref(s))
    dt = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError(f"{impl}_draw failed: {rc}")
    out = dict(targets)
    out.update(fragments=int(s.fragments), primitives_out=int(s.primitives_out), seconds=dt)
    if stream is not None:
        out["stream"] = stream[:min(int(s.stream_len), stream_cap)]
        out["stream_len"] = int(s.stream_len)
    return out


def run_raster_triangles(scene, verts: np.ndarray, impl: str = "oracle", targets: Optional[dict] = None) -> dict:
    """Rasterizer::drawTriangle on screen-space triangles [n,3,7] = {x,y,z,w,a0,a1,a2}."""
    targets = fresh_targets(scene.width, scene.height) if targets is None else targets
    s, keep, _ = _fill(scene, targets, 0)
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3, 7)
    fn = getattr(lib(impl), ("oracle" if impl == "oracle" else "ref") + "_draw_raster_triangles")
    rc = fn(C.byref(s), verts.ctypes.data, verts.shape[0])
    if rc != 0:
        raise RuntimeError(f"{impl}_draw_raster_triangles failed: {rc}")
    out = dict(targets)
    out.update(fragments=int(s.fragments), primitives_out=int(s.primitives_out))
    return out


def run_raster_prims(scene, mode: int, verts: np.ndarray, impl: str = "oracle", targets: Optional[dict] = None) -> dict:
    """Rasterizer::drawPoint / drawLine / drawTriangle on screen-space primitives [n, mode + 1, 7] = {x,y,z,w,a0,a1,a2}."""
    targets = fresh_targets(scene.width, scene.height) if targets is None else targets
    s, keep, _ = _fill(scene, targets, 0)
    verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, mode + 1, 7)
    fn = getattr(lib(impl), ("oracle" if impl == "oracle" else "ref") + "_draw_raster_prims")
    rc = fn(C.byref(s), mode, verts.ctypes.data, verts.shape[0])
    if rc != 0:
        raise RuntimeError(f"{impl}_draw_raster_prims failed: {rc}")
    out = dict(targets)
    out.update(fragments=int(s.fragments), primitives_out=int(s.primitives_out))
    return out


def ref_random_doubles(seed: int, n: int) -> np.ndarray:
    l = lib("ref")
    l.ref_random_doubles.argtypes = [C.c_int, C.c_int64, C.c_void_p]
    out = np.empty(n, dtype=np.float64)
    l.ref_random_doubles(seed, n, out.ctypes.data)
    return out


def ref_box_mvp(eye, fovy=60.0, aspect=4.0 / 3.0, zn=0.1, zf=10.0) -> np.ndarray:
    l = lib("ref")
    l.ref_box_mvp.argtypes = [C.c_float] * 7 + [C.c_void_p]
    out = np.empty(16, dtype=np.float32)
    l.ref_box_mvp(float(eye[0]), float(eye[1]), float(eye[2]), fovy, aspect, zn, zf, out.ctypes.data)
    return out.reshape(4, 4)
