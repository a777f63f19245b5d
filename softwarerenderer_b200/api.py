"""Host-side mirror of the reference's draw-path interface over the C ABI (ctypes).

`Rasterizer` and `VertexProcessor` carry the reference's method names, argument orders, enum
values and defaults (src/renderer/Rasterizer.h:52-141, VertexProcessor.h:56-91).  Shaders are
chosen from the stock pack compiled into libswr_b200.so (custom CRTP shaders are C++: see
include/swr/Renderer.h and examples/).  Everything here only marshals arguments: all computation
happens in the CUDA kernels behind the ABI, and the module cannot be imported without them.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import numpy as np

from . import _lib
from ._lib import SwrError, SwrStats, StockUniforms  # noqa: F401

lib = _lib.load()   # fails loudly when the CUDA extension is missing


class DrawMode:      # VertexProcessor.h:42-46
    Point, Line, Triangle = 0, 1, 2


class CullMode:      # VertexProcessor.h:49-53
    None_, CCW, CW = 0, 1, 2


class RasterMode:    # Rasterizer.h:45-49
    Span, Block, Adaptive = 0, 1, 2


VS_NAMES = {"pos_color": 0, "mvp_color": 1, "mvp_normal_uv": 2}
PS_NAMES = {"flat": 0, "count_id": 1, "gouraud": 2, "gouraud_depth": 3, "vary_dump": 4, "textured": 5, "textured_aniso": 6}
RT_COLOR, RT_DEPTH, RT_COUNT, RT_PRIM_ID, RT_VARY0 = 0, 1, 2, 3, 4
VARY_PLANES = 8


def _pointer_and_bytes(buf):
    """(address, nbytes, keepalive) of a numpy array / torch tensor / raw int address."""
    if buf is None:
        return None, 0, None
    if isinstance(buf, int):
        return buf, 0, None
    if isinstance(buf, np.ndarray):
        if not buf.flags["C_CONTIGUOUS"]:
            raise ValueError("buffer must be C-contiguous")
        return buf.ctypes.data, buf.nbytes, buf
    if hasattr(buf, "data_ptr"):     # torch tensor (host or CUDA)
        if not buf.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return buf.data_ptr(), buf.numel() * buf.element_size(), buf
    raise TypeError(f"unsupported buffer type {type(buf)}")


class Rasterizer:
    """swr::Rasterizer.  Owns the device context (one CUDA stream + scratch)."""

    def __init__(self, device: int = 0):
        ctx = C.c_void_p()
        _lib.check(lib.swr_create(C.byref(ctx), device), "swr_create")
        self._ctx = ctx
        self._keep: Dict[str, object] = {}
        self.setRasterMode(RasterMode.Span)          # Rasterizer.h:67-72
        self.setScissorRect(0, 0, 0, 0)

    def close(self):
        if getattr(self, "_ctx", None):
            lib.swr_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ctx(self):
        return self._ctx

    # ---- reference surface
    def setRasterMode(self, mode: int):
        _lib.check(lib.swr_set_raster_mode(self._ctx, int(mode)), "setRasterMode")

    def setScissorRect(self, x: int, y: int, width: int, height: int):
        _lib.check(lib.swr_set_scissor_rect(self._ctx, x, y, width, height), "setScissorRect")

    def setPixelShader(self, shader):
        """setPixelShader<PS>(): `shader` is a stock pixel shader name or id."""
        kind = PS_NAMES[shader] if isinstance(shader, str) else int(shader)
        ps = lib.swr_stock_pixel_shader(kind)
        if not ps:
            raise SwrError(f"unknown stock pixel shader {shader!r}")
        _lib.check(lib.swr_set_pixel_shader(self._ctx, ps), "setPixelShader")

    def _drawList(self, mode, vertices, indices):
        vp, vbytes, k1 = _pointer_and_bytes(vertices)
        ip, ibytes, k2 = _pointer_and_bytes(indices)
        self._keep["raster_list"] = (k1, k2)
        _lib.check(lib.swr_draw_raster_list(self._ctx, mode, vp, vbytes // 144, ip, ibytes // 4), "draw*List")

    def drawPointList(self, vertices, indices):
        self._drawList(DrawMode.Point, vertices, indices)

    def drawLineList(self, vertices, indices):
        self._drawList(DrawMode.Line, vertices, indices)

    def drawTriangleList(self, vertices, indices):
        """IRasterizer::drawTriangleList on RasterizerVertex records (float32 [n, 36])."""
        self._drawList(DrawMode.Triangle, vertices, indices)

    # ---- additive surface
    def setRenderTarget(self, slot: int, device_ptr: int, pitch_bytes: int, width: int, height: int):
        _lib.check(lib.swr_set_render_target(self._ctx, slot, device_ptr, pitch_bytes, width, height), "setRenderTarget")

    def setUniforms(self, data):
        raw = bytes(data)
        _lib.check(lib.swr_set_uniforms(self._ctx, raw, len(raw)), "setUniforms")

    def setTileSize(self, tile: int):
        _lib.check(lib.swr_set_tile_size(self._ctx, tile), "setTileSize")

    def setIndexNarrowing(self, mode: int):
        """Narrowed upload of big host index arrays (-1 automatic, 0 off, 1 always try); see swr_set_index_narrowing."""
        _lib.check(lib.swr_set_index_narrowing(self._ctx, mode), "setIndexNarrowing")

    def setTileSplit(self, groups: int):
        """Heavy-tile split threshold in listed 32-record groups (-1 automatic, 0 off); see swr_set_tile_split."""
        _lib.check(lib.swr_set_tile_split(self._ctx, groups), "setTileSplit")

    def setTilePartition(self, rank: int, world: int):
        _lib.check(lib.swr_set_tile_partition(self._ctx, rank, world), "setTilePartition")

    def setTileMirrors(self, slot: int, surfaces):
        """Finished tiles of render target `slot` are also stored to these device surfaces (peers' framebuffers)."""
        arr = (C.c_void_p * max(len(surfaces), 1))(*[C.c_void_p(int(p)) for p in surfaces])
        _lib.check(lib.swr_set_tile_mirrors(self._ctx, slot, len(surfaces), arr), "setTileMirrors")

    def ipcOpen(self, handle: bytes, offset: int) -> int:
        out = C.c_void_p()
        _lib.check(lib.swr_ipc_open(self._ctx, handle, offset, C.byref(out)), "ipcOpen")
        return int(out.value)

    def ipcClose(self, ptr: int):
        _lib.check(lib.swr_ipc_close(self._ctx, C.c_void_p(ptr)), "ipcClose")

    def createSharedScratch(self, nbytes: int) -> int:
        """One allocation for the per-pass scratch that the peers of a partition can map (sharded geometry)."""
        out = C.c_void_p()
        _lib.check(lib.swr_shared_scratch_create(self._ctx, nbytes, C.byref(out)), "createSharedScratch")
        return int(out.value)

    def setGeometryShards(self, rank: int, world: int, arenas):
        """arenas[r] = device address of rank r's shared scratch as mapped in this process (own entry ignored)."""
        arr = (C.c_void_p * max(len(arenas), 1))(*[C.c_void_p(int(p) if p else 0) for p in arenas])
        _lib.check(lib.swr_set_geometry_shards(self._ctx, rank, world, arr), "setGeometryShards")

    def peerBarrier(self):
        """Stream-ordered barrier between the ranks of the partition (no NCCL call)."""
        _lib.check(lib.swr_peer_barrier(self._ctx), "peerBarrier")

    def setScratchLimit(self, nbytes: int):
        _lib.check(lib.swr_set_scratch_limit(self._ctx, nbytes), "setScratchLimit")

    def setStream(self, cuda_stream: int):
        """Enqueue on a caller-owned CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); 0 = own stream."""
        _lib.check(lib.swr_set_stream(self._ctx, cuda_stream or None), "setStream")

    def setPipeline(self, enable: bool = True, passes: int = 0, overlap_draws: bool = False):
        """Geometry(k+1) under tiles(k): inside one draw (passes, 0 = automatic) and optionally across draws."""
        _lib.check(lib.swr_set_pipeline(self._ctx, int(enable), int(passes), int(overlap_draws)), "setPipeline")

    def finish(self):
        _lib.check(lib.swr_finish(self._ctx), "finish")

    def stats(self) -> SwrStats:
        s = SwrStats()
        _lib.check(lib.swr_get_stats(self._ctx, C.byref(s)), "get_stats")
        return s

    def resetStats(self):
        _lib.check(lib.swr_reset_stats(self._ctx), "reset_stats")

    def timerBegin(self):
        _lib.check(lib.swr_timer_begin(self._ctx), "timer_begin")

    def timerEnd(self) -> float:
        ms = C.c_float()
        _lib.check(lib.swr_timer_end(self._ctx, C.byref(ms)), "timer_end")
        return float(ms.value)

    def flushL2(self):
        _lib.check(lib.swr_flush_l2(self._ctx), "flush_l2")

    # ---- device memory helpers
    def alloc(self, nbytes: int) -> int:
        p = lib.swr_device_alloc(self._ctx, nbytes)
        if not p:
            raise SwrError(f"device allocation of {nbytes} bytes failed: {lib.swr_last_error().decode()}")
        return p

    def free(self, ptr: int):
        _lib.check(lib.swr_device_free(self._ctx, ptr), "device_free")

    def upload(self, dst: int, src: np.ndarray):
        src = np.ascontiguousarray(src)
        _lib.check(lib.swr_memcpy_h2d(self._ctx, dst, src.ctypes.data, src.nbytes), "memcpy_h2d")
        self._keep["upload"] = src

    def download(self, src: int, out: np.ndarray):
        _lib.check(lib.swr_memcpy_d2h(self._ctx, out.ctypes.data, src, out.nbytes), "memcpy_d2h")

    def fill32(self, dst: int, value: int, count: int):
        _lib.check(lib.swr_memset32(self._ctx, dst, value & 0xFFFFFFFF, count), "memset32")


def ipc_handle(device_ptr: int):
    """(64-byte CUDA IPC handle, offset) of a device pointer, to be opened by a peer process with Rasterizer.ipcOpen."""
    h = C.create_string_buffer(64)
    off = C.c_int64()
    _lib.check(lib.swr_ipc_get_handle(C.c_void_p(int(device_ptr)), h, C.byref(off)), "ipc_handle")
    return h.raw, int(off.value)


class VertexProcessor:
    """swr::VertexProcessor (VertexProcessor.h:56-91); defaults as VertexProcessor.cpp:29-35."""

    def __init__(self, rasterizer: Rasterizer):
        self.setRasterizer(rasterizer)
        self.setCullMode(CullMode.CW)
        self.setDepthRange(0.0, 1.0)
        self._keep: Dict[object, object] = {}

    def setRasterizer(self, rasterizer: Rasterizer):
        assert rasterizer is not None
        self._r = rasterizer

    def setViewport(self, x: int, y: int, width: int, height: int):
        _lib.check(lib.swr_set_viewport(self._r.ctx, x, y, width, height), "setViewport")

    def setDepthRange(self, n: float, f: float):
        _lib.check(lib.swr_set_depth_range(self._r.ctx, n, f), "setDepthRange")

    def setCullMode(self, mode: int):
        _lib.check(lib.swr_set_cull_mode(self._r.ctx, int(mode)), "setCullMode")

    def setVertexShader(self, shader):
        """setVertexShader<VS>(): `shader` is a stock vertex shader name or id."""
        kind = VS_NAMES[shader] if isinstance(shader, str) else int(shader)
        vs = lib.swr_stock_vertex_shader(kind)
        if not vs:
            raise SwrError(f"unknown stock vertex shader {shader!r}")
        _lib.check(lib.swr_set_vertex_shader(self._r.ctx, vs), "setVertexShader")

    def setVertexAttribPointer(self, index: int, stride: int, buffer, nbytes: Optional[int] = None):
        """buffer: numpy array (host, staged per draw), CUDA tensor or device address (used in place)."""
        ptr, n, keep = _pointer_and_bytes(buffer)
        self._keep[index] = keep
        _lib.check(lib.swr_set_vertex_attrib_pointer(self._r.ctx, index, stride, ptr, n if nbytes is None else nbytes),
                   "setVertexAttribPointer")

    def processElements(self, mode: int, count: int, indices):
        """drawElements towards a foreign IRasterizer (IRasterizer.h:56-71): the vertex stage only.  Returns, per batch of
        1024 input primitives, what the reference hands to IRasterizer::draw*List: (vertices float32 [n, 36] =
        RasterizerVertex records in screen space, indices int32 [n] with -1 for dropped primitives)."""
        ptr, n, keep = _pointer_and_bytes(indices)
        out = []

        def emit(user, draw_mode, verts, nverts, idx, nidx):
            v = np.ctypeslib.as_array(C.cast(verts, C.POINTER(C.c_float)), shape=(nverts, 36)).copy() if nverts else np.zeros((0, 36), np.float32)
            i = np.ctypeslib.as_array(C.cast(idx, C.POINTER(C.c_int32)), shape=(nidx,)).copy() if nidx else np.zeros(0, np.int32)
            out.append((v, i))
        cb = _lib.STREAM_OUT_FN(emit)
        _lib.check(lib.swr_process_elements(self._r.ctx, int(mode), int(count), ptr, cb, None), "processElements")
        return out

    def drawElements(self, mode: int, count: int, indices, wait: bool = True):
        """drawElements(mode, count, indices); like the reference, complete on return unless wait=False."""
        ptr, n, keep = _pointer_and_bytes(indices)
        self._keep["indices"] = keep
        _lib.check(lib.swr_draw_elements(self._r.ctx, int(mode), int(count), ptr), "drawElements")
        if wait:
            self._r.finish()


class RenderTargets:
    """The twelve 32-bit surfaces the stock pixel shaders use, in device memory."""
    NAMES = ("color", "depth", "count", "prim_id")

    def __init__(self, r: Rasterizer, width: int, height: int):
        self.r, self.width, self.height = r, width, height
        self.n = width * height
        self.pitch = width * 4
        self.base = r.alloc(self.n * 4 * _lib.MAX_RENDER_TARGETS)
        for s in range(_lib.MAX_RENDER_TARGETS):
            r.setRenderTarget(s, self.ptr(s), self.pitch, width, height)
        self.clear()

    def ptr(self, slot: int) -> int:
        return self.base + slot * self.n * 4

    def clear(self):
        """color 0, depth 1.0f, count 0, prim_id 0xFFFFFFFF, vary 0 (same as oracle.pyoracle.fresh_targets)."""
        r = self.r
        r.fill32(self.ptr(RT_COLOR), 0, self.n)
        r.fill32(self.ptr(RT_DEPTH), 0x3F800000, self.n)
        r.fill32(self.ptr(RT_COUNT), 0, self.n)
        r.fill32(self.ptr(RT_PRIM_ID), 0xFFFFFFFF, self.n)
        r.fill32(self.ptr(RT_VARY0), 0, self.n * VARY_PLANES)

    def download(self) -> dict:
        r = self.r
        out = {
            "color": np.empty(self.n, dtype=np.uint32),
            "depth": np.empty(self.n, dtype=np.float32),
            "count": np.empty(self.n, dtype=np.uint32),
            "prim_id": np.empty(self.n, dtype=np.uint32),
            "vary": np.empty(self.n * VARY_PLANES, dtype=np.float32),
        }
        r.download(self.ptr(RT_COLOR), out["color"])
        r.download(self.ptr(RT_DEPTH), out["depth"])
        r.download(self.ptr(RT_COUNT), out["count"])
        r.download(self.ptr(RT_PRIM_ID), out["prim_id"])
        r.download(self.ptr(RT_VARY0), out["vary"])
        r.finish()
        return out

    def free(self):
        if self.base:
            self.r.free(self.base)
            self.base = 0


class SceneRenderer:
    """Draws softwarerenderer_b200.scenes.Scene objects through the reference-shaped API."""

    def __init__(self, width: int, height: int, device: int = 0, tile_size: int = 0):
        self.r = Rasterizer(device)
        self.v = VertexProcessor(self.r)
        self.targets = RenderTargets(self.r, width, height)
        if tile_size:
            self.r.setTileSize(tile_size)
        self._tex = None
        self._tex_key = None

    def close(self):
        if self._tex:
            self.r.free(self._tex)
            self._tex = None
        self.targets.free()
        self.r.close()

    def _uniforms(self, scene):
        u = StockUniforms()
        u.mvp = (C.c_float * 16)(*[float(x) for x in scene.mvp.reshape(-1)])
        if scene.texture is not None:
            tex = np.ascontiguousarray(scene.texture, dtype=np.uint32)
            key = (tex.ctypes.data, tex.shape)
            if self._tex_key != key:
                from .scenes import build_mip_chain
                if self._tex:
                    self.r.free(self._tex)
                self._mips = build_mip_chain(tex)                     # level 0 = the texture itself
                chain = np.concatenate([m.reshape(-1) for m in self._mips])
                chain[:tex.size] = tex.reshape(-1)                    # nearest-fetch shader reads the raw texels
                self._tex = self.r.alloc(chain.nbytes)
                self.r.upload(self._tex, chain)
                self._tex_key = key
            u.texture = self._tex
            u.tex_h, u.tex_w = tex.shape
            off = 0
            for i, m in enumerate(self._mips[:_lib.MAX_MIP_LEVELS]):
                u.mip[i] = self._tex + off * 4
                u.mip_h[i], u.mip_w[i] = m.shape
                off += m.size
            u.mip_levels = min(len(self._mips), _lib.MAX_MIP_LEVELS)
            u.max_anisotropy = 8                                      # Texture.h:14 default
        return u

    def set_state(self, scene):
        r, v = self.r, self.v
        assert scene.width == self.targets.width and scene.height == self.targets.height
        r.setRasterMode(scene.raster_mode)
        r.setScissorRect(*scene.scissor)
        r.setPixelShader(scene.ps)
        v.setViewport(*scene.viewport)
        v.setDepthRange(*scene.depth_range)
        v.setCullMode(scene.cull_mode)
        v.setVertexShader(scene.vs)
        r.setUniforms(self._uniforms(scene))

    def draw(self, scene, vertices=None, indices=None, wait: bool = True):
        """One drawElements call.  vertices / indices default to the scene's host arrays (staged by
        the library); pass device addresses or CUDA tensors to draw from resident buffers."""
        self.set_state(scene)
        vb = scene.vertices if vertices is None else vertices
        ib = scene.indices if indices is None else indices
        self.v.setVertexAttribPointer(0, scene.stride, vb, scene.vertices.nbytes)
        self.v.drawElements(scene.draw_mode, int(scene.indices.size), ib, wait=wait)

    def render(self, scene, clear: bool = True) -> dict:
        """Clear, draw, read everything back: same dictionary as oracle.pyoracle.run()."""
        if clear:
            self.targets.clear()
        self.r.resetStats()
        self.draw(scene)
        out = self.targets.download()
        st = self.r.stats()
        out.update(fragments=int(st.fragments), stats=st)
        return out
