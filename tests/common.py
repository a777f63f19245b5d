"""Scene lists shared by the CPU and GPU parity tests."""
import json
import os

import numpy as np

from softwarerenderer_b200 import scenes as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BUFFERS = ("color", "depth", "count", "prim_id", "vary")


def box_mesh():
    d = np.load(os.path.join(GOLDEN, "box_mesh.npz"))
    return d["vertices"], d["indices"], d["mvp_theta0"], d["mvp_theta05"], d["mvp_theta2"], d["mvp_near"]


def box_texture():
    """data/box.png as 0x00RRGGBB words (fixture made by tests/golden/make_golden.py)."""
    return np.load(os.path.join(GOLDEN, "box_texture.npz"))["texture"].astype(np.uint32)


def heavy_clip_mvp():
    """v*3 - 1.5 on x, y, z as a matrix (heavy 6-plane clipping)."""
    m = np.eye(4, dtype=np.float32)
    m[0, 0] = m[1, 1] = m[2, 2] = 3
    m[:3, 3] = -1.5
    return m


def parity_scenes(ntri=2048, small=True):
    """(label, scene) pairs covering every draw mode, raster mode, cull mode and stock shader."""
    out = []
    for mode in (S.RASTER_SPAN, S.RASTER_BLOCK, S.RASTER_ADAPTIVE):
        for ps in (S.PS_COUNT_ID, S.PS_GOURAUD):
            out.append((f"vptest_m{mode}_ps{ps}", S.vertex_processor_test(mode, ps)))
    b0 = S.config_c0(ps=S.PS_COUNT_ID, ntri=ntri)
    h = b0.replace(vs=S.VS_MVP_COLOR, mvp=heavy_clip_mvp())
    for sc, nm in ((b0, "bench"), (h, "heavy")):
        for mode in (S.RASTER_SPAN, S.RASTER_BLOCK, S.RASTER_ADAPTIVE):
            for cull in (S.CULL_NONE, S.CULL_CCW, S.CULL_CW):
                out.append((f"{nm}_m{mode}_c{cull}", sc.replace(raster_mode=mode, cull_mode=cull)))
        edges = S.triangle_edges(sc.indices)
        out.append((f"{nm}_points", sc.replace(draw_mode=S.DRAW_POINT)))
        out.append((f"{nm}_lines", sc.replace(draw_mode=S.DRAW_LINE, indices=edges)))
        out.append((f"{nm}_flat_span", sc.replace(ps=S.PS_FLAT)))
        out.append((f"{nm}_depth_block", sc.replace(raster_mode=S.RASTER_BLOCK, ps=S.PS_GOURAUD_DEPTH)))
        out.append((f"{nm}_depth_span", sc.replace(raster_mode=S.RASTER_SPAN, ps=S.PS_GOURAUD_DEPTH)))
        out.append((f"{nm}_lines_depth", sc.replace(draw_mode=S.DRAW_LINE, indices=edges, ps=S.PS_GOURAUD_DEPTH)))
    bv, bi, m0, m05, m2, mnear = box_mesh()
    tex = S.checker_texture()
    for th, m in ((0.0, m0), (0.5, m05), (2.0, m2)):
        for mode in (S.RASTER_SPAN, S.RASTER_BLOCK, S.RASTER_ADAPTIVE):
            for ps in (S.PS_TEXTURED, S.PS_VARY_DUMP, S.PS_COUNT_ID):
                out.append((f"box_th{th}_m{mode}_ps{ps}", S.config_c1(bv, bi, tex, th, raster_mode=mode, ps=ps, mvp=m)))
    for mode in (S.RASTER_SPAN, S.RASTER_BLOCK):
        near = S.config_c1(bv, bi, tex, 0, raster_mode=mode, ps=S.PS_VARY_DUMP, mvp=mnear).replace(cull_mode=S.CULL_NONE)
        out.append((f"boxnear_m{mode}_vary", near))
        out.append((f"boxnear_m{mode}_tex", near.replace(ps=S.PS_TEXTURED)))
        out.append((f"boxnear_m{mode}_lines", near.replace(draw_mode=S.DRAW_LINE, indices=S.triangle_edges(bi))))
        out.append((f"boxnear_m{mode}_points", near.replace(draw_mode=S.DRAW_POINT)))
    out.append(("c2_small", S.config_c2(100, 50, 480, 270)))
    out.append(("c2_small_span", S.config_c2(100, 50, 480, 270, raster_mode=S.RASTER_SPAN)))
    out.append(("c3_small", S.config_c3(250, 200, 480, 270)))
    out.append(("c3_small_id", S.config_c3(250, 200, 480, 270, ps=S.PS_COUNT_ID)))
    out.append(("c4_lines_small", S.config_c4(100, 50, 480, 270)))
    out.append(("c4_points_small", S.config_c4(100, 50, 480, 270, draw_mode=S.DRAW_POINT)))
    # Box.cpp's own pixel shader: perspective derivatives + the reference's Texture::sample (Texture.h)
    for th, m in ((0.0, m0), (0.5, m05), (2.0, m2)):
        out.append((f"box_th{th}_aniso_span", S.config_c1(bv, bi, tex, th, raster_mode=S.RASTER_SPAN, ps=S.PS_TEXTURED_ANISO, mvp=m)))
    out.append(("boxnear_aniso_block", S.config_c1(bv, bi, tex, 0, raster_mode=S.RASTER_BLOCK, ps=S.PS_TEXTURED_ANISO, mvp=mnear).replace(cull_mode=S.CULL_NONE)))
    # BASELINE.json configs[0] as the reference ships it: box.obj + box.png, 640x480, Span, Box.cpp's shaders
    out.append(("c1_box_png_aniso_span", S.config_c1(bv, bi, box_texture(), 0.5, raster_mode=S.RASTER_SPAN, ps=S.PS_TEXTURED_ANISO, mvp=m05)))
    out.append(("c5_small_aniso", S.config_c5(100, 80, 3, 480, 270, ps=S.PS_TEXTURED_ANISO)))
    out.append(("c5_small", S.config_c5(100, 80, 3, 480, 270)))
    out.append(("c5_small_vary", S.config_c5(100, 80, 3, 480, 270, ps=S.PS_VARY_DUMP)))
    return out


def diff_buffers(got, want, keys=BUFFERS):
    """{buffer: number of differing 32-bit words}; bit-exact comparison (floats as their bits)."""
    bad = {}
    for k in keys:
        d = int((got[k].view(np.uint32) != want[k].view(np.uint32)).sum())
        if d:
            bad[k] = d
    return bad


def max_channel_diff(a, b):
    """Largest per-channel difference between two 0x00RRGGBB buffers, and how many pixels differ."""
    a, b = a.astype(np.int64), b.astype(np.int64)
    d = 0
    for sh in (16, 8, 0):
        d = np.maximum(d, np.abs(((a >> sh) & 255) - ((b >> sh) & 255)))
    return int(d.max()), int((d > 0).sum())


def known_answers():
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        return json.load(f)
