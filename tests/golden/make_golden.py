"""Regenerates tests/golden/* from the UNMODIFIED reference build (oracle/_ref/libswr_ref.so).

Run in the development container only (it needs /root/reference):
    python tests/golden/make_golden.py
Outputs (committed):
    box_texture.npz     data/box.png decoded to 0x00RRGGBB words
    box_mesh.npz        data/box.obj through our OBJ loader + the Box.cpp cameras computed by the
                        reference's own vector_math.h (ref_box_mvp)
    random0_prefix.npy  first 256 values of the reference's Random(0).NextDouble()
    known_answers.json  per scene: fragments, covered pixels, primitives handed to the rasterizer and
                        a CRC32 of every output buffer, all produced by the reference renderer
"""
import json
import math
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from oracle import pyoracle as O  # noqa: E402
from softwarerenderer_b200 import scenes as S  # noqa: E402


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).view(np.uint8).tobytes()) & 0xFFFFFFFF


def summarize(out):
    d = {"fragments": out["fragments"], "primitives_out": out["primitives_out"]}
    for k in ("color", "depth", "count", "prim_id", "vary"):
        d["crc_" + k] = crc(out[k])
    d["covered"] = int((out["count"] > 0).sum())
    return d


def main():
    O.build()
    assert O.have_ref(), "reference build missing"
    # ---- box mesh + cameras
    bv, bi = S.load_obj(open("/root/reference/data/box.obj").read())
    assert bv.shape == (24, 8) and bi.shape == (36,)

    def eye(th):
        th = np.float32(th)
        return (np.float32(5.0) * np.float32(math.cos(th)), 2.0, np.float32(5.0) * np.float32(math.sin(th)))
    np.savez(os.path.join(HERE, "box_mesh.npz"), vertices=bv, indices=bi,
             mvp_theta0=O.ref_box_mvp(eye(0.0)), mvp_theta05=O.ref_box_mvp(eye(0.5)), mvp_theta2=O.ref_box_mvp(eye(2.0)),
             mvp_near=O.ref_box_mvp((1.2, 0.3, 0.4)))
    np.save(os.path.join(HERE, "random0_prefix.npy"), O.ref_random_doubles(0, 256))
    # data/box.png decoded to the 0x00RRGGBB words Texture.h expects (alpha is 255 everywhere)
    from PIL import Image
    rgba = np.asarray(Image.open("/root/reference/data/box.png").convert("RGBA"), dtype=np.uint32)
    np.savez_compressed(os.path.join(HERE, "box_texture.npz"), texture=(rgba[..., 0] << 16) | (rgba[..., 1] << 8) | rgba[..., 2])

    import common  # tests/common.py (needs box_mesh.npz)
    answers = {}
    for label, scene in common.parity_scenes():
        answers[label] = summarize(O.run(scene, "ref"))
        print(label, answers[label]["fragments"])
    # the reference's own benchmark at full size
    full = S.config_c0(ps=S.PS_COUNT_ID)
    for mode, name in ((0, "span"), (1, "block"), (2, "adaptive")):
        answers[f"benchmark_full_{name}"] = summarize(O.run(full.replace(raster_mode=mode), "ref"))
    answers["benchmark_full_points"] = summarize(O.run(full.replace(draw_mode=S.DRAW_POINT), "ref"))
    answers["benchmark_full_lines"] = summarize(O.run(full.replace(draw_mode=S.DRAW_LINE, indices=S.triangle_edges(full.indices)), "ref"))
    sc = S.Scene("rt", np.zeros((1, 6), np.float32), np.zeros(3, np.int32), 640, 480, ps=S.PS_COUNT_ID)
    for mode, name in ((0, "span"), (1, "block"), (2, "adaptive")):
        out = O.run_raster_triangles(sc.replace(raster_mode=mode), S.rasterizer_test_triangle(), "ref")
        answers[f"rasterizer_test_{name}"] = summarize(out)
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(answers, f, indent=1, sort_keys=True)
    print("wrote", len(answers), "known answers")


if __name__ == "__main__":
    main()
