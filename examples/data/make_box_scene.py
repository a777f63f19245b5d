"""Writes examples/data/box_scene.bin from the committed fixtures tests/golden/box_mesh.npz (the box mesh in the order
ObjData::toVertexArray produces, the Box.cpp:190-195 camera at 0 / 0.5 / 2.0 rad) and tests/golden/box_texture.npz
(data/box.png as 0x00RRGGBB words), both made by tests/golden/make_golden.py.
Layout: int32 {magic 'SBOX', vertices, indices, tex_w, tex_h}, float32 vertices[n][8] {pos3, normal3, uv2},
int32 indices[], 3 x float32 mvp[16] (row-major), uint32 texels[tex_h][tex_w]."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
d = np.load(os.path.join(ROOT, "tests", "golden", "box_mesh.npz"))
t = np.load(os.path.join(ROOT, "tests", "golden", "box_texture.npz"))["texture"].astype(np.uint32)
with open(os.path.join(ROOT, "examples", "data", "box_scene.bin"), "wb") as f:
    f.write(np.array([0x584F4253, d["vertices"].shape[0], d["indices"].shape[0], t.shape[1], t.shape[0]], dtype=np.int32).tobytes())
    f.write(d["vertices"].astype(np.float32).tobytes())
    f.write(d["indices"].astype(np.int32).tobytes())
    for k in ("mvp_theta0", "mvp_theta05", "mvp_theta2"):
        f.write(d[k].astype(np.float32).tobytes())
    f.write(t.tobytes())
