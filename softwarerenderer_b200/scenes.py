"""Synthetic inputs for the draw path: meshes, cameras and the five BASELINE.json configs.

Everything here is host-side input generation in numpy (no rendering).  The same bytes are
handed to the CUDA path, to the oracle and to the reference build, so parity never depends
on how an input was produced.

Reference shapes followed:
  * Benchmark.cpp:53-95   -- 40 960 random triangles from Random(0), 6 NextDouble() per vertex
  * Random.cpp:7-50       -- the Knuth subtractive generator behind Random::NextDouble()
  * ObjData.cpp:69-200    -- OBJ parsing, fan triangulation and (v, n, t) de-duplication order
  * Box.cpp:190-195       -- perspective(60, 4/3, 0.1, 10) * lookat(orbiting eye, 0, +Y)
  * SURVEY.md section 8(d) -- configs C1..C5
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

# Reference enum values (VertexProcessor.h:42-53, Rasterizer.h:45-49).
DRAW_POINT, DRAW_LINE, DRAW_TRIANGLE = 0, 1, 2
CULL_NONE, CULL_CCW, CULL_CW = 0, 1, 2
RASTER_SPAN, RASTER_BLOCK, RASTER_ADAPTIVE = 0, 1, 2

# Stock shader ids (include/swr_b200.h, oracle/swr_scene.h).
VS_POS_COLOR, VS_MVP_COLOR, VS_MVP_NORMAL_UV = 0, 1, 2
PS_FLAT, PS_COUNT_ID, PS_GOURAUD, PS_GOURAUD_DEPTH, PS_VARY_DUMP, PS_TEXTURED, PS_TEXTURED_ANISO = 0, 1, 2, 3, 4, 5, 6

ORDINAL_STRIDE = 10240
VARY_PLANES = 8

IDENTITY = np.eye(4, dtype=np.float32)


@dataclass
class Scene:
    """One draw call: geometry + fixed-function state + stock shader ids + target size."""
    name: str
    vertices: np.ndarray            # float32 [num_vertices, stride/4]
    indices: np.ndarray             # int32 [index_count]
    width: int
    height: int
    draw_mode: int = DRAW_TRIANGLE
    cull_mode: int = CULL_CW        # reference default (VertexProcessor.cpp:32)
    raster_mode: int = RASTER_SPAN  # reference default (Rasterizer.h:69)
    vs: int = VS_POS_COLOR
    ps: int = PS_FLAT
    viewport: Optional[Tuple[int, int, int, int]] = None
    scissor: Optional[Tuple[int, int, int, int]] = None
    depth_range: Tuple[float, float] = (0.0, 1.0)
    mvp: np.ndarray = field(default_factory=lambda: IDENTITY.copy())
    texture: Optional[np.ndarray] = None   # uint32 [tex_h, tex_w]

    def __post_init__(self):
        self.vertices = np.ascontiguousarray(self.vertices, dtype=np.float32)
        self.indices = np.ascontiguousarray(self.indices, dtype=np.int32)
        self.mvp = np.ascontiguousarray(self.mvp, dtype=np.float32).reshape(4, 4)
        if self.viewport is None:
            self.viewport = (0, 0, self.width, self.height)
        if self.scissor is None:
            self.scissor = (0, 0, self.width, self.height)

    @property
    def stride(self) -> int:
        return int(self.vertices.shape[1] * 4)

    @property
    def num_vertices(self) -> int:
        return int(self.vertices.shape[0])

    @property
    def num_primitives(self) -> int:
        return int(self.indices.size // (self.draw_mode + 1))

    def replace(self, **kw) -> "Scene":
        d = dict(self.__dict__)
        d.update(kw)
        return Scene(**d)


# --------------------------------------------------------------------------- Random.cpp
_MBIG = 2147483647
_MSEED = 161803398


def dotnet_random_doubles(seed: int, n: int) -> np.ndarray:
    """First n values of Random(seed).NextDouble() (Random.cpp:7-50,129-132)."""
    sa = [0] * 56
    sub = _MBIG if seed == -2147483648 else abs(seed)
    mj = _MSEED - sub
    sa[55] = mj
    mk = 1
    for i in range(1, 55):
        ii = (21 * i) % 55
        sa[ii] = mk
        mk = mj - mk
        if mk < 0:
            mk += _MBIG
        mj = sa[ii]
    for _ in range(4):
        for i in range(1, 56):
            sa[i] -= sa[1 + (i + 30) % 55]
            if sa[i] < 0:
                sa[i] += _MBIG
    inext, inextp = 0, 21
    out = np.empty(n, dtype=np.float64)
    scale = 1.0 / _MBIG
    for k in range(n):
        inext += 1
        if inext >= 56:
            inext = 1
        inextp += 1
        if inextp >= 56:
            inextp = 1
        r = sa[inext] - sa[inextp]
        if r == _MBIG:
            r -= 1
        if r < 0:
            r += _MBIG
        sa[inext] = r
        out[k] = r * scale
    return out


# --------------------------------------------------------------------------- cameras
def perspective(fovy_deg: float, aspect: float, znear: float, zfar: float) -> np.ndarray:
    """Row-major OpenGL-style projection (same convention as vector_math.h:737-759)."""
    f32 = np.float32
    rad = f32(fovy_deg) / f32(2) * f32(math.pi / 180.0)
    cot = f32(math.cos(rad)) / f32(math.sin(rad))
    dz = f32(zfar) - f32(znear)
    m = np.zeros((4, 4), dtype=np.float32)
    m[0, 0] = cot / f32(aspect)
    m[1, 1] = cot
    m[2, 2] = -(f32(zfar) + f32(znear)) / dz
    m[2, 3] = f32(-2) * f32(znear) * f32(zfar) / dz
    m[3, 2] = f32(-1)
    return m


def lookat(eye, center=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0)) -> np.ndarray:
    eye = np.asarray(eye, dtype=np.float32)
    center = np.asarray(center, dtype=np.float32)
    up = np.asarray(up, dtype=np.float32)
    fwd = center - eye
    fwd = fwd / np.float32(np.linalg.norm(fwd))
    side = np.cross(fwd, up)
    side = side / np.float32(np.linalg.norm(side))
    up2 = np.cross(side, fwd)
    m = np.eye(4, dtype=np.float32)
    m[0, :3] = side
    m[1, :3] = up2
    m[2, :3] = -fwd
    t = np.eye(4, dtype=np.float32)
    t[:3, 3] = -eye
    return (m @ t).astype(np.float32)


def rotation_x(angle: float) -> np.ndarray:
    c, s = np.float32(math.cos(angle)), np.float32(math.sin(angle))
    m = np.eye(4, dtype=np.float32)
    m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    return m


def translation(x: float, y: float, z: float) -> np.ndarray:
    m = np.eye(4, dtype=np.float32)
    m[:3, 3] = (x, y, z)
    return m


# --------------------------------------------------------------------------- meshes
def benchmark_mesh(ntri: int = 40960, seed: int = 0):
    """Benchmark.cpp:53-95: unshared vertices {x,y,z,r,g,b} in NextDouble() order."""
    d = dotnet_random_doubles(seed, ntri * 3 * 6).astype(np.float32)
    vertices = d.reshape(ntri * 3, 6)
    indices = np.arange(ntri * 3, dtype=np.int32)
    return vertices, indices


def grid_indices(nx: int, ny: int) -> np.ndarray:
    """Two triangles per cell, (a,b,d) and (a,d,c) with a=(i,j) b=(i+1,j) c=(i,j+1) d=(i+1,j+1);
    row-major cell order, so consecutive primitives are spatial neighbours."""
    j, i = np.meshgrid(np.arange(ny, dtype=np.int64), np.arange(nx, dtype=np.int64), indexing="ij")
    a = j * (nx + 1) + i
    b = a + 1
    c = a + (nx + 1)
    d = c + 1
    tri = np.stack([a, b, d, a, d, c], axis=-1)
    return tri.reshape(-1).astype(np.int32)


def grid_positions(nx: int, ny: int, half_w: float, half_h: float):
    xs = np.linspace(-half_w, half_w, nx + 1, dtype=np.float64)
    ys = np.linspace(-half_h, half_h, ny + 1, dtype=np.float64)
    y, x = np.meshgrid(ys, xs, indexing="ij")
    return x, y


def colors(n: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).random((n, 3), dtype=np.float32)


def triangle_edges(indices: np.ndarray) -> np.ndarray:
    """a-b, b-c, c-a per triangle."""
    t = indices.reshape(-1, 3)
    e = np.stack([t[:, 0], t[:, 1], t[:, 1], t[:, 2], t[:, 2], t[:, 0]], axis=-1)
    return e.reshape(-1).astype(np.int32)


def load_obj(text: str):
    """OBJ -> ({pos3, normal3, uv2} float32 [n, 8], int32 indices) in the order
    ObjData::toVertexArray produces (ObjData.cpp:139-200): fan triangulation, first-seen
    numbering of distinct (v, n, t) triples."""
    vs, ns, ts = [(0.0, 0.0, 0.0)], [(0.0, 0.0, 0.0)], [(0.0, 0.0)]
    faces = []
    for line in text.splitlines():
        p = line.split()
        if not p or p[0].startswith("#"):
            continue
        if p[0] == "v":
            vs.append(tuple(float(q) for q in p[1:4]))
        elif p[0] == "vn":
            ns.append(tuple(float(q) for q in p[1:4]))
        elif p[0] == "vt":
            ts.append(tuple(float(q) for q in p[1:3]))
        elif p[0] == "f":
            face = []
            for w in p[1:]:
                q = (w.split("/") + ["", ""])[:3]
                face.append((int(q[0] or 0), int(q[2] or 0), int(q[1] or 0)))  # (v, n, t)
            faces.append(face)
    index_of, vdata, idata = {}, [], []

    def add(ref):
        if ref not in index_of:
            index_of[ref] = len(index_of)
            vdata.append(vs[ref[0]] + ns[ref[1]] + ts[ref[2]])
        return index_of[ref]

    for face in faces:
        i1 = add(face[0])
        for k in range(2, len(face)):
            i2 = add(face[k - 1])
            i3 = add(face[k])
            idata += [i1, i2, i3]
    return np.asarray(vdata, dtype=np.float32), np.asarray(idata, dtype=np.int32)


def checker_texture(size: int = 256, seed: int = 7) -> np.ndarray:
    """A deterministic 0x00RRGGBB texture with per-texel detail (stand-in for data/box.png)."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 1 << 24, size=(size, size), dtype=np.uint32)
    yy, xx = np.meshgrid(np.arange(size), np.arange(size), indexing="ij")
    check = (((xx >> 4) ^ (yy >> 4)) & 1).astype(np.uint32)
    return (base & np.uint32(0x003F3F3F)) | (check * np.uint32(0x00C0C0C0))


def build_mip_chain(base: np.ndarray):
    """Texture::generateMipmaps (Texture.h:220-293): 2x2 box filter per channel with >> 2, down to 1x1.
    Returns the list of uint32 [h, w] levels (level 0 = base)."""
    lv = (np.ascontiguousarray(base, dtype=np.uint32) & np.uint32(0x00FFFFFF))
    levels = [lv]
    while lv.shape[0] > 1 or lv.shape[1] > 1:
        h, w = lv.shape
        nh, nw = max(1, h // 2), max(1, w // 2)
        ys, xs = np.arange(nh) * 2, np.arange(nw) * 2
        y1 = np.where(ys + 1 < h, ys + 1, -1)
        x1 = np.where(xs + 1 < w, xs + 1, -1)
        p00 = lv[np.ix_(ys, xs)]
        p10 = np.where((x1 >= 0)[None, :], lv[np.ix_(ys, np.maximum(x1, 0))], p00)
        p01 = np.where((y1 >= 0)[:, None], lv[np.ix_(np.maximum(y1, 0), xs)], p00)
        p11 = np.where((y1 >= 0)[:, None] & (x1 >= 0)[None, :], lv[np.ix_(np.maximum(y1, 0), np.maximum(x1, 0))], p00)
        out = np.zeros((nh, nw), dtype=np.uint32)
        for shift in (16, 8, 0):
            ch = sum(((p >> np.uint32(shift)) & np.uint32(0xFF)).astype(np.uint32) for p in (p00, p10, p01, p11)) >> np.uint32(2)
            out |= (ch & np.uint32(0xFF)) << np.uint32(shift)
        levels.append(out)
        lv = out
    return levels


# --------------------------------------------------------------------------- configs
def config_c0(width: int = 640, height: int = 480, ntri: int = 40960, raster_mode: int = RASTER_SPAN,
              ps: int = PS_FLAT, seed: int = 0) -> Scene:
    """The reference's own benchmark (Benchmark.cpp:66-105): fill-bound, big triangles."""
    v, i = benchmark_mesh(ntri, seed)
    return Scene(f"c0_benchmark_{width}x{height}", v, i, width, height, cull_mode=CULL_NONE,
                 raster_mode=raster_mode, vs=VS_POS_COLOR, ps=ps)


def box_camera(theta: float, aspect: float = 4.0 / 3.0) -> np.ndarray:
    eye = (np.float32(5.0) * np.float32(math.cos(np.float32(theta))), np.float32(2.0),
           np.float32(5.0) * np.float32(math.sin(np.float32(theta))))
    return (perspective(60.0, aspect, 0.1, 10.0) @ lookat(eye)).astype(np.float32)


def config_c1(box_vertices: np.ndarray, box_indices: np.ndarray, texture: np.ndarray, theta: float = 0.0,
              raster_mode: int = RASTER_SPAN, ps: int = PS_TEXTURED, mvp: Optional[np.ndarray] = None) -> Scene:
    """data/box.obj textured box, 640x480, Span, 3 affine + 2 perspective varyings."""
    return Scene(f"c1_box_theta{theta}", box_vertices, box_indices, 640, 480, cull_mode=CULL_CW,
                 raster_mode=raster_mode, vs=VS_MVP_NORMAL_UV, ps=ps,
                 mvp=box_camera(theta) if mvp is None else mvp, texture=texture)


def _grid_vertices_color(x, y, z, seed):
    n = x.size
    v = np.empty((n, 6), dtype=np.float32)
    v[:, 0] = x.reshape(-1)
    v[:, 1] = y.reshape(-1)
    v[:, 2] = z.reshape(-1)
    v[:, 3:6] = colors(n, seed)
    return v


def config_c2(nx: int = 1000, ny: int = 500, width: int = 1920, height: int = 1080,
              raster_mode: int = RASTER_BLOCK, ps: int = PS_GOURAUD_DEPTH) -> Scene:
    """1M-triangle grid, Gouraud + depth test, perspective camera."""
    aspect = width / height
    x, y = grid_positions(nx, ny, 1.0, 1.0)
    z = 0.25 * np.sin(7.0 * x) * np.cos(5.0 * y)
    # fill the view at the z=0 plane seen from eye z=2 with fov 60
    half_h = 2.0 * math.tan(math.radians(30.0)) * 0.97
    v = _grid_vertices_color(x * (half_h * aspect), y * half_h, z, seed=1)
    mvp = perspective(60.0, aspect, 0.1, 10.0) @ lookat((0.0, 0.0, 2.0))
    return Scene(f"c2_grid_{nx}x{ny}_{width}x{height}", v, grid_indices(nx, ny), width, height,
                 cull_mode=CULL_CW, raster_mode=raster_mode, vs=VS_MVP_COLOR, ps=ps, mvp=mvp)


def config_c3(nx: int = 2500, ny: int = 2000, width: int = 3840, height: int = 2160,
              raster_mode: int = RASTER_BLOCK, ps: int = PS_GOURAUD) -> Scene:
    """10M tiny triangles: a sheet folded in x (about half of the triangles end up clockwise and are
    removed by CullMode::CW) and tilted about the x axis through the near plane (near-plane clipping,
    plus some +-X / +-Y clipping at the sides)."""
    aspect = width / height
    x, y = grid_positions(nx, ny, 1.0, 1.0)
    folds = 24.0
    xf = x + (6.0 / (folds * math.pi)) * np.sin(folds * math.pi * x)     # dx'/dx = 1 + 6 cos() < 0 on ~45%
    z = 0.02 * np.cos(folds * math.pi * x)
    v = _grid_vertices_color(xf * 0.8 * aspect, y * 1.6, z, seed=3)
    # Tilted about x so that the near rows pass behind the eye: ~11% of the triangles carry the -Z
    # clip bit, ~17% each +-X, ~32% -Y; ~45% of the unclipped ones are clockwise; ~1/3 survive.
    model = rotation_x(-1.2)
    view = lookat((0.0, 0.3, 1.1), (0.0, 0.0, -0.3))
    mvp = perspective(60.0, aspect, 0.1, 10.0) @ view @ model
    return Scene(f"c3_tiny_{nx}x{ny}_{width}x{height}", v, grid_indices(nx, ny), width, height,
                 cull_mode=CULL_CW, raster_mode=raster_mode, vs=VS_MVP_COLOR, ps=ps, mvp=mvp)


def config_c4(nx: int = 1000, ny: int = 500, width: int = 3840, height: int = 2160,
              draw_mode: int = DRAW_LINE, ps: int = PS_GOURAUD) -> Scene:
    """Wireframe / point cloud of C2's mesh at 4K with the camera pulled in (LineClipper path)."""
    base = config_c2(nx, ny, width, height)
    aspect = width / height
    mvp = perspective(60.0, aspect, 0.1, 10.0) @ lookat((0.3, 0.1, 1.45))
    idx = triangle_edges(base.indices) if draw_mode == DRAW_LINE else base.indices
    kind = "lines" if draw_mode == DRAW_LINE else "points"
    return Scene(f"c4_{kind}_{nx}x{ny}_{width}x{height}", base.vertices, idx, width, height,
                 draw_mode=draw_mode, cull_mode=CULL_CW, raster_mode=RASTER_BLOCK, vs=VS_MVP_COLOR, ps=ps, mvp=mvp)


def config_c5(nx: int = 2500, ny: int = 2000, layers: int = 5, width: int = 7680, height: int = 4320,
              raster_mode: int = RASTER_BLOCK, ps: int = PS_TEXTURED, texture: Optional[np.ndarray] = None) -> Scene:
    """layers x (nx x ny x 2) perspective-textured triangles, painter's order = submission order."""
    aspect = width / height
    x, y = grid_positions(nx, ny, 1.0, 1.0)
    per = (nx + 1) * (ny + 1)
    v = np.empty((layers * per, 8), dtype=np.float32)
    idx = []
    base_idx = grid_indices(nx, ny)
    half_h = 2.0 * math.tan(math.radians(30.0))
    for l in range(layers):
        s = 0.55 + 0.1 * l                       # later layers are larger and nearer: real overdraw
        z = 0.15 * np.sin(3.0 * x + l) * np.cos(4.0 * y - l) + 0.12 * l - 0.3
        o = v[l * per:(l + 1) * per]
        o[:, 0] = (x * (half_h * aspect * s)).reshape(-1)
        o[:, 1] = (y * (half_h * s)).reshape(-1)
        o[:, 2] = z.reshape(-1)
        o[:, 3:5] = 0.0
        o[:, 5] = 1.0
        o[:, 6] = ((x * 0.5 + 0.5) * 64.0).reshape(-1)
        o[:, 7] = ((y * 0.5 + 0.5) * 64.0).reshape(-1)
        idx.append(base_idx + np.int32(l * per))
    mvp = perspective(60.0, aspect, 0.1, 10.0) @ lookat((0.25, 0.15, 2.0))
    tex = checker_texture() if texture is None else texture
    return Scene(f"c5_layers{layers}_{nx}x{ny}_{width}x{height}", v, np.concatenate(idx), width, height,
                 cull_mode=CULL_CW, raster_mode=raster_mode, vs=VS_MVP_NORMAL_UV, ps=ps, mvp=mvp, texture=tex)


def config_fill(nx: int = 96, ny: int = 54, layers: int = 2, width: int = 3840, height: int = 2160,
                raster_mode: int = RASTER_BLOCK, ps: int = PS_FLAT) -> Scene:
    """Fill-bound, low overdraw (SURVEY.md 8(d) "worked bounds"): `layers` full-screen sheets of nx x ny x 2 large
    triangles (40 pixels across at 4K), every pixel shaded `layers` times with a 4-byte store -- the regime in which
    the frame-buffer traffic, not the per-primitive work, is what a draw moves."""
    x, y = grid_positions(nx, ny, 1.0, 1.0)
    per = (nx + 1) * (ny + 1)
    v = np.empty((layers * per, 6), dtype=np.float32)
    idx = []
    base_idx = grid_indices(nx, ny)
    for l in range(layers):
        o = v[l * per:(l + 1) * per]
        o[:, 0] = x.reshape(-1)
        o[:, 1] = y.reshape(-1)
        o[:, 2] = 0.5 - 0.1 * l
        o[:, 3:6] = colors(per, 40 + l)
        idx.append(base_idx + np.int32(l * per))
    return Scene(f"fill_{layers}x{nx}x{ny}_{width}x{height}", v, np.concatenate(idx), width, height,
                 cull_mode=CULL_NONE, raster_mode=raster_mode, vs=VS_POS_COLOR, ps=ps)


def rasterizer_test_triangle() -> np.ndarray:
    """RasterizerTest.cpp:60-80 as 1 triangle x 3 vertices x {x,y,z,w,a0,a1,a2} (z=0, w=1)."""
    return np.asarray([[320, 100, 0, 1, 1, 0, 0], [480, 200, 0, 1, 0, 1, 0], [120, 300, 0, 1, 0, 0, 1]],
                      dtype=np.float32)


def vertex_processor_test(raster_mode: int = RASTER_SPAN, ps: int = PS_GOURAUD) -> Scene:
    """VertexProcessorTest.cpp:78-121: one clip-space triangle clipped by +-X, offset viewport."""
    v = np.asarray([[0.0, 0.5, 0.0, 1, 0, 0], [-1.5, -0.5, 0.0, 0, 1, 0], [1.5, -0.5, 0.0, 0, 0, 1]], dtype=np.float32)
    return Scene("vertex_processor_test", v, np.asarray([0, 1, 2], dtype=np.int32), 640, 480,
                 cull_mode=CULL_NONE, raster_mode=raster_mode, vs=VS_POS_COLOR, ps=ps,
                 viewport=(100, 100, 440, 280), scissor=(0, 0, 640, 480))
