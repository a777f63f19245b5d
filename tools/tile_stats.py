"""Per-tile time / work distribution of one draw (debug).  usage: python tools/tile_stats.py c3 [tile]"""
import ctypes as C
import os
import sys

import numpy as np

os.environ.setdefault("SWR_LIB_VARIANT", "_stats")   # make -C softwarerenderer_b200/csrc VARIANT=_stats EXTRA=-DSWR_TILE_STATS=1

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
from gpu_time import SCENES  # noqa: E402
from softwarerenderer_b200 import _lib  # noqa: E402
from softwarerenderer_b200.api import SceneRenderer  # noqa: E402

lib = _lib.load()
scene = SCENES[sys.argv[1]]()
tile = int(sys.argv[2]) if len(sys.argv) > 2 else 0
sr = SceneRenderer(scene.width, scene.height, tile_size=tile)
vb = sr.r.alloc(scene.vertices.nbytes)
ib = sr.r.alloc(scene.indices.nbytes)
sr.r.upload(vb, scene.vertices)
sr.r.upload(ib, scene.indices)
sr.draw(scene, vertices=vb, indices=ib)
lib.swr_debug_enable_tile_stats(sr.r.ctx, 1)
sr.draw(scene, vertices=vb, indices=ib)
st = sr.r.stats()
T = st.last_tile_size
tx, ty = (scene.width + T - 1) // T, (scene.height + T - 1) // T
buf = np.zeros((tx * ty, 16), dtype=np.uint32)
n = lib.swr_debug_read_tile_stats(sr.r.ctx, buf.ctypes.data, tx * ty)
start, dur, prims, frags = [buf[:, i].astype(np.int64) for i in range(4)]
cA0, cA, cB = [buf[:, i].astype(np.int64) * 16 / 1.965e3 for i in (4, 5, 6)]   # us at 1965 MHz
t0 = start[dur > 0].min()
rel = (start - t0) & 0xFFFFFFFF
print(f"{scene.name} tile{T}: {n} tiles, kernel {st.last_tile_ms:.3f} ms; busy tiles {(prims > 0).sum()}")
print(f"  duration us: mean {dur.mean() / 1e3:.1f}  p50 {np.percentile(dur, 50) / 1e3:.1f}  p90 {np.percentile(dur, 90) / 1e3:.1f}  p99 {np.percentile(dur, 99) / 1e3:.1f}  max {dur.max() / 1e3:.1f}")
print(f"  sum of durations {dur.sum() / 1e6:.2f} ms -> / (148 SMs x 3 CTAs) = {dur.sum() / 1e6 / 444:.3f} ms ideal")
print(f"  prims per tile: mean {prims.mean():.0f} max {prims.max()};  frags per tile: mean {frags.mean():.0f} max {frags.max()}")
print(f"  thread-0 phase time summed over tiles (ms): pre-test {cA0.sum() / 1e3:.1f}  coverage {cA.sum() / 1e3:.1f}  shading {cB.sum() / 1e3:.1f}  "
      f"binning+rest {(dur.sum() / 1e3 - cA0.sum() - cA.sum() - cB.sum()) / 1e3:.1f}  (total {dur.sum() / 1e6:.1f})")
cPre, cF3, cF12 = [buf[:, i].astype(np.int32).astype(np.int64) * 16 / 1.965e3 for i in (8, 9, 10)]
print(f"  of binning+rest (ms): flush prologue {cPre.sum() / 1e3:.1f}  F3 records {cF3.sum() / 1e3:.1f}  F1+F2 chunks/groups {cF12.sum() / 1e3:.1f};  "
      f"records tested {buf[:, 11].astype(np.int64).sum()}  groups tested {buf[:, 12].astype(np.int64).sum()}  queued {prims.sum()}")
w7 = buf[:, 7].astype(np.int64)
nfl = w7
print(f"  flushes {nfl.sum()}")
order = np.argsort(-dur)[:12]
for i in order:
    print(f"    tile ({i % tx},{i // tx}) start +{rel[i] / 1e3:8.1f} us dur {dur[i] / 1e3:8.1f} us prims {prims[i]:7d} frags {frags[i]:7d}  A0 {cA0[i]:6.1f} A {cA[i]:6.1f} B {cB[i]:6.1f} pre {cPre[i]:5.1f} F3 {cF3[i]:6.1f} F12 {cF12[i]:6.1f} us flushes {nfl[i]} tested {buf[i, 11]}/{buf[i, 12]}")
end = (rel + dur)
print(f"  last tile ends at +{end.max() / 1e3:.1f} us; tiles starting after 50% of that: {(rel > end.max() / 2).sum()}")
rows = dur.reshape(ty, tx).sum(axis=1) / 1e3
print("  per tile-row sum of durations (us):", " ".join(f"{r:.0f}" for r in rows))
