"""Minimal driver for ncu: draws one of tools/gpu_time.py's scenes three times with resident inputs.
usage: ncu ... python tools/prof_run.py c3 [tile]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gpu_time import SCENES  # noqa: E402
from softwarerenderer_b200.api import SceneRenderer  # noqa: E402

scene = SCENES[sys.argv[1]]()
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sr = SceneRenderer(scene.width, scene.height, tile_size=tile)
vb = sr.r.alloc(scene.vertices.nbytes)
ib = sr.r.alloc(scene.indices.nbytes)
sr.r.upload(vb, scene.vertices)
sr.r.upload(ib, scene.indices)
for _ in range(3):
    sr.draw(scene, vertices=vb, indices=ib)
print("done", sr.r.stats().fragments, "tile", sr.r.stats().last_tile_size)
