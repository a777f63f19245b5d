"""Small draws of every kind for compute-sanitizer (memcheck / racecheck / synccheck).
usage: compute-sanitizer --tool racecheck python tools/sanitize_run.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from softwarerenderer_b200 import scenes as S  # noqa: E402
from softwarerenderer_b200.api import SceneRenderer  # noqa: E402

cases = [
    S.config_c3(120, 100, 320, 200),                                   # tiny triangles, clipping, extras
    S.config_c3(120, 100, 320, 200, raster_mode=S.RASTER_SPAN),
    S.config_c0(320, 200, ntri=300, ps=S.PS_GOURAUD_DEPTH, raster_mode=S.RASTER_BLOCK),   # big triangles: pre-test + dense path
    S.config_c4(60, 40, 320, 200),                                     # lines
    S.config_c4(60, 40, 320, 200, draw_mode=S.DRAW_POINT),
    S.config_c5(40, 30, 2, 320, 200, ps=S.PS_TEXTURED_ANISO),
]
for tile in (32, 64):
    for sc in cases:
        sr = SceneRenderer(sc.width, sc.height, tile_size=tile)
        out = sr.render(sc)
        print(tile, sc.name, out["fragments"], flush=True)
        sr.close()
# heavy-tile split: every busy tile shaded by four quadrant CTAs
for tile in (32, 64):
    for sc in (cases[0], cases[2], cases[3]):
        sr = SceneRenderer(sc.width, sc.height, tile_size=tile)
        sr.r.setTileSplit(1)
        out = sr.render(sc)
        print("split", tile, sc.name, out["fragments"], flush=True)
        sr.close()
# sharded geometry: two contexts on this GPU push records into each other's scratch, ordered by the flag barrier
sc = S.config_c3(120, 100, 320, 200)
srs = [SceneRenderer(sc.width, sc.height) for _ in range(2)]
for sr in srs:
    sr.draw(sc)
arenas = [sr.r.createSharedScratch(64 << 20) for sr in srs]
for rank, sr in enumerate(srs):
    sr.r.setGeometryShards(rank, 2, arenas)
    sr.targets.clear()
    sr.r.resetStats()
for sr in srs:
    sr.r.finish()
for sr in srs:
    sr.draw(sc, wait=False)
for sr in srs:
    sr.r.peerBarrier()
frags = 0
for sr in srs:
    sr.r.finish()
    frags += sr.r.stats().fragments
print("sharded x2", sc.name, frags, flush=True)
for sr in srs:
    sr.r.setGeometryShards(0, 1, [])
    sr.close()
# the vertex stage alone (foreign IRasterizer)
from softwarerenderer_b200.api import Rasterizer, VertexProcessor  # noqa: E402
sr = SceneRenderer(sc.width, sc.height)
sr.set_state(sc)
sr.v.setVertexAttribPointer(0, sc.stride, sc.vertices, sc.vertices.nbytes)
print("stream-out batches", len(sr.v.processElements(sc.draw_mode, int(sc.indices.size), sc.indices)), flush=True)
sr.close()
print("sanitize run done")
