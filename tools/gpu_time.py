"""GPU-box timing of the big configs with device-resident inputs (CUDA events on the context stream).
usage: python tools/gpu_time.py [c0 c0b c2 c3 c0_4k ...] [--tile 32|64] [--check]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from softwarerenderer_b200 import scenes as S  # noqa: E402
from softwarerenderer_b200.api import SceneRenderer  # noqa: E402

SCENES = {
    "c0": lambda: S.config_c0(ps=S.PS_FLAT, raster_mode=S.RASTER_SPAN),
    "c0b": lambda: S.config_c0(ps=S.PS_FLAT, raster_mode=S.RASTER_BLOCK),
    "c0g": lambda: S.config_c0(ps=S.PS_GOURAUD_DEPTH, raster_mode=S.RASTER_BLOCK),
    "c0_4k": lambda: S.config_c0(3840, 2160, ps=S.PS_FLAT, raster_mode=S.RASTER_BLOCK),
    "c2": S.config_c2,
    "c3": S.config_c3,
    "c3s": lambda: S.config_c3(raster_mode=S.RASTER_SPAN),
    "c4l": S.config_c4,
    "c4p": lambda: S.config_c4(draw_mode=S.DRAW_POINT),
    "c5": S.config_c5,
    "c5a": lambda: S.config_c5(ps=S.PS_TEXTURED_ANISO),
    "fill": S.config_fill,
}


def timed(scene, tile, reps=5, check=False):
    sr = SceneRenderer(scene.width, scene.height, tile_size=tile)
    vb = sr.r.alloc(scene.vertices.nbytes)
    ib = sr.r.alloc(scene.indices.nbytes)
    sr.r.upload(vb, scene.vertices)
    sr.r.upload(ib, scene.indices)
    sr.r.finish()
    ts = []
    for i in range(reps):
        sr.targets.clear()
        sr.r.resetStats()
        sr.r.flushL2()
        sr.r.timerBegin()
        sr.draw(scene, vertices=vb, indices=ib, wait=False)
        ms = sr.r.timerEnd()
        st = sr.r.stats()
        ts.append((ms, st.last_geometry_ms, st.last_tile_ms, st.fragments))
    # K back-to-back draws (throughput), optionally with the next draw's geometry under this draw's tiles
    for overlap in (0, 1):
        sr.r.setPipeline(True, 0, bool(overlap))
        for _ in range(2):
            sr.draw(scene, vertices=vb, indices=ib, wait=False)
        sr.r.finish()
        K = 10
        sr.r.timerBegin()
        for _ in range(K):
            sr.draw(scene, vertices=vb, indices=ib, wait=False)
        ms = sr.r.timerEnd()
        sr.r.finish()
        print(f"   back-to-back x{K} overlap_draws={overlap}: {ms / K:.3f} ms/draw", flush=True)
    sr.r.setPipeline(False, 0, False)
    best = min(ts)
    print(f"TIME {scene.name:36s} tile{st.last_tile_size} prims {scene.num_primitives:>9d} frags {best[3]:>11d} total {best[0]:8.3f}ms geom {best[1]:7.3f} tile {best[2]:8.3f} "
          f"-> {scene.num_primitives / best[0] / 1e6:6.2f} Gprim/s {best[3] / best[0] / 1e6:7.2f} Gfrag/s", flush=True)
    if check:
        from oracle import pyoracle as O
        import common
        O.build()
        got = sr.targets.download()
        want = O.run(scene, "ref" if O.have_ref() else "oracle")
        bad = common.diff_buffers(got, want, ("color", "depth"))
        print("   CHECK", "OK" if not bad and best[3] == want["fragments"] else f"MISMATCH {bad} frags {best[3]} vs {want['fragments']}", flush=True)
    sr.r.free(vb)
    sr.r.free(ib)
    sr.close()


if __name__ == "__main__":
    argv = list(sys.argv[1:])
    tiles = [32, 64]
    if "--tile" in argv:
        i = argv.index("--tile")
        tiles = [int(argv[i + 1])]
        del argv[i:i + 2]
    args = [a for a in argv if not a.startswith("--")]
    for name in args or ["c0", "c0b", "c2", "c3"]:
        sc = SCENES[name]()
        for tile in tiles:
            timed(sc, tile, check="--check" in sys.argv)
