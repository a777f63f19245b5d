"""CPU: the C-ABI library loads, exports every symbol include/swr_b200.h declares, and refuses to
run without a GPU (no CPU fallback).  Also the host-side logic of the Python mirror."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "swr_b200.h")).read()
    return sorted(set(re.findall(r"SWR_API\s+[\w\s\*]+?\b(swr_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from softwarerenderer_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 35
    bound = {n for n, _, _ in _lib.SYMBOLS}
    for n in names:
        assert hasattr(lib, n), f"libswr_b200.so does not export {n}"
        assert n in bound, f"softwarerenderer_b200/_lib.py does not bind {n}"
    assert lib.swr_abi_version() == 1


def test_stock_shader_descriptors():
    from softwarerenderer_b200 import _lib
    lib = _lib.load()
    for k in range(3):
        assert lib.swr_stock_vertex_shader(k)
    for k in range(6):
        assert lib.swr_stock_pixel_shader(k)
    assert not lib.swr_stock_vertex_shader(99) and not lib.swr_stock_pixel_shader(99)


def test_owned_tile_count_partitions_the_screen():
    from softwarerenderer_b200 import _lib
    lib = _lib.load()
    for (w, h, t) in ((3840, 2160, 64), (1920, 1080, 32), (640, 480, 32), (7680, 4320, 64)):
        total = ((w + t - 1) // t) * ((h + t - 1) // t)
        for world in (1, 2, 4, 8):
            counts = [lib.swr_owned_tile_count(w, h, t, r, world) for r in range(world)]
            assert sum(counts) == total
            assert max(counts) - min(counts) <= max(2, total // 50), (w, h, t, world, counts)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from softwarerenderer_b200 import api
    with pytest.raises(api.SwrError, match="no CUDA device"):
        api.Rasterizer()


def test_product_never_imports_oracle():
    """The product path must not route through the checkers."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "softwarerenderer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle|#include\s+[\"<].*oracle", text, re.M):
                    bad.append(f)
    for base, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            text = open(os.path.join(base, f), errors="ignore").read()
            if re.search(r"#include\s+[\"<].*oracle/", text):
                bad.append(f)
    assert not bad, bad


def test_scene_generators_shapes():
    from softwarerenderer_b200 import scenes as S
    s = S.config_c2(10, 5, 64, 32)
    assert s.vertices.shape == (66, 6) and s.indices.size == 10 * 5 * 6 and s.stride == 24
    s = S.config_c5(4, 3, 2, 64, 32)
    assert s.vertices.shape == (2 * 20, 8) and s.indices.size == 2 * 4 * 3 * 6 and s.stride == 32
    e = S.triangle_edges(np.arange(6, dtype=np.int32))
    assert e.tolist() == [0, 1, 1, 2, 2, 0, 3, 4, 4, 5, 5, 3]
    v, i = S.benchmark_mesh(4)
    assert v.shape == (12, 6) and i.tolist() == list(range(12))
