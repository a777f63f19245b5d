"""Tile phase of every rank of a sort-first partition, measured on ONE GPU (the partition only selects which tiles a
context shades, so rank r of N can be timed without the other GPUs; geometry runs replicated here).
usage: python tools/rank_emul.py [c3|c5|c2 ...] [--worlds 1,2,4,8]
Per (world, tile size, split threshold): max and mean over the ranks of the bin + tile kernel time (best of 5 draws)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gpu_time import SCENES  # noqa: E402
from softwarerenderer_b200.api import SceneRenderer  # noqa: E402

argv = sys.argv[1:]
worlds = [1, 2, 4, 8]
if "--worlds" in argv:
    i = argv.index("--worlds")
    worlds = [int(x) for x in argv[i + 1].split(",")]
    del argv[i:i + 2]
CONFIGS = [(32, 0), (64, 0), (64, 128), (64, 256), (64, 512)]
if "--configs" in argv:
    i = argv.index("--configs")
    CONFIGS = [tuple(int(v) for v in x.split(":")) for x in argv[i + 1].split(",")]
    del argv[i:i + 2]

for name in argv or ["c3"]:
    scene = SCENES[name]()
    sr = SceneRenderer(scene.width, scene.height)
    vb = sr.r.alloc(scene.vertices.nbytes)
    ib = sr.r.alloc(scene.indices.nbytes)
    sr.r.upload(vb, scene.vertices)
    sr.r.upload(ib, scene.indices)
    sr.r.finish()
    for world in worlds:
        for tile, split in CONFIGS:
            if world == 1 and split and False:
                continue
            sr.r.setTileSize(tile)
            sr.r.setTileSplit(split)
            per_rank, frags = [], 0
            for rank in range(world):
                sr.r.setTilePartition(rank, world)
                best = 1e9
                for _ in range(5):
                    sr.r.resetStats()
                    sr.draw(scene, vertices=vb, indices=ib, wait=True)
                    st = sr.r.stats()
                    best = min(best, st.last_tile_ms)
                frags += int(st.fragments)
                per_rank.append(best)
            print(f"EMUL {name} world {world} tile {tile} split {split:4d}: tile phase max {max(per_rank):.3f} mean {np.mean(per_rank):.3f} ms  "
                  f"(per rank {' '.join(f'{x:.3f}' for x in per_rank)})  frags {frags}", flush=True)
    sr.r.free(vb)
    sr.r.free(ib)
    sr.close()
