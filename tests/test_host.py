"""CPU: host-side helpers next to the draw path -- the OBJ loader (ObjData.cpp:69-200 order), the headless presenter
(Box.cpp:204's SDL_UpdateWindowSurface as a file) and the scene generators' bookkeeping."""
import os

import numpy as np
import pytest

import common
from softwarerenderer_b200 import present, scenes as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

QUAD_OBJ = """
# one quad and one triangle sharing an edge
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 2 0 0
vn 0 0 1
vt 0 0
vt 1 0
vt 1 1
vt 0 1
f 1/1/1 2/2/1 3/3/1 4/4/1
f 2/2/1 5/1/1 3/3/1
"""


def test_load_obj_fan_triangulation_and_first_seen_numbering():
    """ObjData::toVertexArray (ObjData.cpp:139-200): polygons become fans around their first corner, distinct
    (v, n, t) triples are numbered in first-seen order."""
    v, i = S.load_obj(QUAD_OBJ)
    assert i.tolist() == [0, 1, 2, 0, 2, 3, 1, 4, 2]
    assert v.shape == (5, 8)
    assert v[4].tolist() == [2, 0, 0, 0, 0, 1, 0, 0]          # position 5, normal 1, texcoord 1
    assert v[2].tolist() == [1, 1, 0, 0, 0, 1, 1, 1]


def test_load_obj_reproduces_the_box_fixture():
    """data/box.obj through load_obj == the committed fixture tests/golden/box_mesh.npz (made from the reference's own
    ObjData by tests/golden/make_golden.py).  Needs the reference tree for the .obj text."""
    path = "/root/reference/data/box.obj"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    v, i = S.load_obj(open(path).read())
    bv, bi = common.box_mesh()[:2]
    assert np.array_equal(i, bi) and np.array_equal(v.view(np.uint32), bv.view(np.uint32))


def test_presenter_writes_the_colour_buffer(tmp_path):
    w, h = 7, 5
    color = (np.arange(w * h, dtype=np.uint32) * 0x010203) & 0xFFFFFF
    path = tmp_path / "frame.ppm"
    present.write_ppm(str(path), color, w, h)
    raw = path.read_bytes()
    head = b"P6\n7 5\n255\n"
    assert raw.startswith(head) and len(raw) == len(head) + w * h * 3
    rgb = np.frombuffer(raw[len(head):], dtype=np.uint8).reshape(h, w, 3)
    assert np.array_equal(rgb, present.to_rgb8(color, w, h))
    assert rgb[0, 1].tolist() == [1, 2, 3]
    present.write_image(str(tmp_path / "frame.png"), color, w, h)       # PNG with Pillow, else PPM next to it
    assert (tmp_path / "frame.png").exists() or (tmp_path / "frame.ppm").exists()


def test_scene_generators_have_the_named_shapes():
    """BASELINE.json's configs as scenes.py builds them: counts, strides, modes (no rendering)."""
    c2 = S.config_c2(100, 50, 480, 270)
    assert c2.num_primitives == 100 * 50 * 2 and c2.stride == 24 and c2.raster_mode == S.RASTER_BLOCK and c2.ps == S.PS_GOURAUD_DEPTH
    c4 = S.config_c4(20, 10, 480, 270)
    assert c4.draw_mode == S.DRAW_LINE and c4.num_primitives == 20 * 10 * 2 * 3
    c5 = S.config_c5(20, 10, 3, 480, 270)
    assert c5.num_primitives == 3 * 20 * 10 * 2 and c5.stride == 32 and c5.texture.shape == (256, 256)
    e = S.triangle_edges(np.array([0, 1, 2], np.int32))
    assert e.tolist() == [0, 1, 1, 2, 2, 0]
    d = S.dotnet_random_doubles(0, 3)
    assert abs(d[0] - 0.7262432699679598) < 1e-15            # Random(0).NextDouble(), Random.cpp:7-50


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the arm the driver runs next to the GPU arm): ONE JSON line on stdout with the
    contract's keys, measured on the reference build (or the oracle port) -- here with a one-second budget."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c0", "--steps", "1",
                          "--warmup", "0", "--ref-budget", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "shaded_fragments_per_s" and d["unit"] == "fragments/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert "Benchmark.cpp" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] == d["value"] and "primitives" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0
