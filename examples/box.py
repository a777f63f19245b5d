#!/usr/bin/env python
"""examples/box.py -- the reference's Box demo (src/examples/Box.cpp) headless: data/box.obj with data/box.png,
640x480, RasterMode::Span, CullMode::CW, Box.cpp's vertex and pixel shaders (perspective-correct UVs, anisotropic
mip-mapped sampling), a few frames of the orbiting camera written as images instead of an SDL window.

    python examples/box.py [frames] [outdir]        (needs a B200; inputs come from tests/golden/)
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from softwarerenderer_b200 import scenes as S  # noqa: E402
from softwarerenderer_b200.api import SceneRenderer  # noqa: E402
from softwarerenderer_b200.present import write_image  # noqa: E402


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    outdir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out")
    os.makedirs(outdir, exist_ok=True)
    mesh = np.load(os.path.join(ROOT, "tests", "golden", "box_mesh.npz"))
    tex = np.load(os.path.join(ROOT, "tests", "golden", "box_texture.npz"))["texture"].astype(np.uint32)
    sr = SceneRenderer(640, 480)
    for k in range(frames):
        angle = 0.5 * k                                               # Box.cpp:187-190: angularSpeed * t
        scene = S.config_c1(mesh["vertices"], mesh["indices"], tex, angle, raster_mode=S.RASTER_SPAN, ps=S.PS_TEXTURED_ANISO)
        out = sr.render(scene)                                        # clears (SDL_FillRect) and draws
        path = os.path.join(outdir, f"box_{k:02d}.png")
        write_image(path, out["color"], 640, 480)
        print(f"frame {k}: angle {angle:.2f} rad, {out['fragments']} fragments -> {path}")
    sr.close()


if __name__ == "__main__":
    main()
