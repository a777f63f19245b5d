"""GPU-box diagnostic: renders a ladder of scenes, diffs against the oracle, prints where they differ,
then times the big configs.  Usage (under gpurun): python tools/gpu_diag.py [quick]"""
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import common  # noqa: E402
from oracle import pyoracle as O  # noqa: E402
from softwarerenderer_b200 import scenes as S  # noqa: E402
from softwarerenderer_b200.api import SceneRenderer  # noqa: E402

O.build()
renderers = {}


def get(w, h, tile=0):
    k = (w, h, tile)
    if k not in renderers:
        renderers[k] = SceneRenderer(w, h, tile_size=tile)
    return renderers[k]


def describe(scene, got, want):
    w = scene.width
    for k in common.BUFFERS:
        g, o = got[k].view(np.uint32), want[k].view(np.uint32)
        idx = np.nonzero(g != o)[0]
        if idx.size == 0:
            continue
        print(f"    {k}: {idx.size} words differ; first:", end="")
        for i in idx[:6]:
            p = i % (scene.width * scene.height)
            print(f" ({p % w},{p // w} plane{i // (scene.width * scene.height)}) gpu={g[i]:#x} ora={o[i]:#x};", end="")
        print()
        if k == "count":
            ys, xs = np.divmod(idx, w)
            print(f"      bbox of count diffs: x[{xs.min()},{xs.max()}] y[{ys.min()},{ys.max()}]; gpu sum {int(got['count'].sum())} oracle sum {int(want['count'].sum())}")


def run_one(label, scene, tile=0):
    try:
        t0 = time.time()
        got = get(scene.width, scene.height, tile).render(scene)
        t1 = time.time()
        want = O.run(scene, "oracle")
        bad = common.diff_buffers(got, want)
        ok = not bad and got["fragments"] == want["fragments"]
        st = got["stats"]
        print(f"{'OK  ' if ok else 'FAIL'} {label:32s} tile{st.last_tile_size} frags gpu {got['fragments']:>10d} oracle {want['fragments']:>10d} "
              f"geom {st.last_geometry_ms:.3f}ms tile {st.last_tile_ms:.3f}ms wall {t1 - t0:.2f}s {bad if bad else ''}", flush=True)
        if not ok:
            describe(scene, got, want)
        return ok
    except Exception:
        print(f"EXC  {label}")
        traceback.print_exc()
        return False


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    nfail = 0
    scenes = common.parity_scenes(ntri=512 if quick else 2048)
    for label, scene in scenes:
        nfail += 0 if run_one(label, scene) else 1
        if nfail > 12:
            print("too many failures, stopping the ladder")
            break
    for tile in (32, 64):
        nfail += 0 if run_one(f"c3_small_tile{tile}", S.config_c3(250, 200, 480, 270, ps=S.PS_COUNT_ID), tile) else 1
    print(f"ladder done: {nfail} failures", flush=True)

    # timing of the big configs (device-resident inputs)
    def timed(scene, tile, reps=5):
        sr = get(scene.width, scene.height, tile)
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
        sr.r.free(vb)
        sr.r.free(ib)
        best = min(ts)
        print(f"TIME {scene.name:40s} tile{tile} prims {scene.num_primitives:>9d} frags {best[3]:>10d} total {best[0]:.3f}ms geom {best[1]:.3f} tile {best[2]:.3f} "
              f"-> {scene.num_primitives / best[0] / 1e6:.2f} Gprim/s {best[3] / best[0] / 1e6:.2f} Gfrag/s (all: {[round(t[0], 3) for t in ts]})", flush=True)

    try:
        for tile in (32, 64):
            timed(S.config_c0(ps=S.PS_FLAT, raster_mode=S.RASTER_BLOCK), tile)
            timed(S.config_c0(ps=S.PS_FLAT, raster_mode=S.RASTER_SPAN), tile)
            timed(S.config_c2(), tile)
            if not quick:
                timed(S.config_c3(), tile)
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
