"""CPU: pins the oracle (oracle/swr_oracle.c) -- against the golden vectors produced by the
unmodified reference build (tests/golden/known_answers.json, made by tests/golden/make_golden.py)
and, when oracle/_ref/libswr_ref.so is present, against that build itself buffer for buffer."""
import zlib

import numpy as np
import pytest

import common
from softwarerenderer_b200 import scenes as S


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a).view(np.uint8).tobytes()) & 0xFFFFFFFF


def check_against_golden(out, want, label):
    assert out["fragments"] == want["fragments"], label
    assert out["primitives_out"] == want["primitives_out"], label
    assert int((out["count"] > 0).sum()) == want["covered"], label
    for k in common.BUFFERS:
        assert crc(out[k]) == want["crc_" + k], f"{label}: buffer {k} differs from the reference's"


SCENES = common.parity_scenes()


@pytest.mark.parametrize("label,scene", SCENES, ids=[l for l, _ in SCENES])
def test_oracle_matches_reference_golden(oracle, label, scene):
    check_against_golden(oracle.run(scene, "oracle"), common.known_answers()[label], label)


@pytest.mark.parametrize("name,mode", [("span", 0), ("block", 1), ("adaptive", 2)])
def test_oracle_reference_benchmark_full(oracle, name, mode):
    """Benchmark.cpp's own workload (40 960 triangles, Random(0)); SURVEY.md section 4 known answers."""
    scene = S.config_c0(ps=S.PS_COUNT_ID, raster_mode=mode)
    out = oracle.run(scene, "oracle")
    check_against_golden(out, common.known_answers()[f"benchmark_full_{name}"], name)
    assert out["fragments"] == {"span": 240235639, "block": 240235776, "adaptive": 240235758}[name]


def test_oracle_benchmark_points_lines(oracle):
    full = S.config_c0(ps=S.PS_COUNT_ID)
    ka = common.known_answers()
    out = oracle.run(full.replace(draw_mode=S.DRAW_POINT), "oracle")
    check_against_golden(out, ka["benchmark_full_points"], "points")
    assert (out["fragments"], int((out["count"] > 0).sum())) == (122880, 61245)
    out = oracle.run(full.replace(draw_mode=S.DRAW_LINE, indices=S.triangle_edges(full.indices)), "oracle")
    check_against_golden(out, ka["benchmark_full_lines"], "lines")
    assert (out["fragments"], int((out["count"] > 0).sum())) == (16239234, 76768)


@pytest.mark.parametrize("name,mode,frags", [("span", 0, 25900), ("block", 1, 26100), ("adaptive", 2, 25900)])
def test_oracle_rasterizer_test_triangle(oracle, name, mode, frags):
    """RasterizerTest.cpp:60-80 through Rasterizer::drawTriangle."""
    sc = S.Scene("rt", np.zeros((1, 6), np.float32), np.zeros(3, np.int32), 640, 480, ps=S.PS_COUNT_ID, raster_mode=mode)
    out = oracle.run_raster_triangles(sc, S.rasterizer_test_triangle(), "oracle")
    assert out["fragments"] == frags
    check_against_golden(out, common.known_answers()[f"rasterizer_test_{name}"], name)


def test_dotnet_random_matches_reference_prefix(golden_dir):
    import os
    want = np.load(os.path.join(golden_dir, "random0_prefix.npy"))
    assert np.array_equal(S.dotnet_random_doubles(0, want.size), want)


def test_oracle_equals_reference_build_live(oracle):
    """Buffer-for-buffer, stream-for-stream equality with the reference compiled in place."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libswr_ref.so not present (no /root/reference on this box)")
    for label, scene in SCENES[::3]:
        a = oracle.run(scene, "oracle", stream_cap=20000)
        b = oracle.run(scene, "ref", stream_cap=20000)
        assert a["fragments"] == b["fragments"] and a["primitives_out"] == b["primitives_out"], label
        assert not common.diff_buffers(a, b), label
        assert a["stream_len"] == b["stream_len"], label
        assert np.array_equal(a["stream"].view(np.uint32), b["stream"].view(np.uint32)), label


def test_empty_and_ragged_inputs(oracle):
    """count == 0, and a count that ends exactly on / just past a 1024-primitive batch."""
    base = S.config_c0(ps=S.PS_COUNT_ID, ntri=1025)
    out = oracle.run(base.replace(indices=base.indices[:0]), "oracle")
    assert out["fragments"] == 0 and out["primitives_out"] == 0
    a = oracle.run(base.replace(indices=base.indices[:3 * 1024]), "oracle")
    b = oracle.run(base, "oracle")
    assert b["primitives_out"] == a["primitives_out"] + 1
    assert b["prim_id"][b["count"] > 0].max() == S.ORDINAL_STRIDE   # the 1025th triangle opens batch 1
