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


def test_index_narrowing_host_packer_round_trip():
    """hostpack.cpp (the host side of the narrowed index upload): base per block of 4096 + 16-bit offsets reproduce the
    indices exactly; a block that spans more than 65535 vertices makes the slice fall back to the plain copy."""
    from softwarerenderer_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(7)

    def pack(idx):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        off = np.zeros(max(1, idx.size), dtype=np.uint16)
        bases = np.full((idx.size + 4095) // 4096 + 1, -7, dtype=np.int32)
        rc = lib.swr_debug_pack_indices16(idx.ctypes.data, idx.size, off.ctypes.data, bases.ctypes.data)
        return rc, off[:idx.size], bases

    rc, _, _ = pack(np.arange(10))
    if rc == -1:
        pytest.skip("no AVX2 on this CPU: the library never narrows")
    for n in (1, 7, 8, 15, 16, 17, 4095, 4096, 4097, 12288, 100003, 1 << 20):
        # a mesh-like stream: a slowly moving window of 60000 vertices, far above 2^16 in absolute value
        centre = 3_000_000 + (np.arange(n) // 3) * 2
        idx = (centre + rng.integers(0, 60000, n)).astype(np.int32)
        rc, off, bases = pack(idx)
        assert rc == 1, n
        nb = (n + 4095) // 4096
        assert bases[nb] == -7                                          # nothing written past the last block
        want_base = np.array([idx[b * 4096:(b + 1) * 4096].min() for b in range(nb)], dtype=np.int32)
        assert np.array_equal(bases[:nb], want_base)
        assert np.array_equal(np.repeat(bases[:nb], 4096)[:n] + off.astype(np.int32), idx)
    # the limit: a span of exactly 65535 still fits, 65536 does not (in any position of any block)
    idx = np.full(9000, 1000, dtype=np.int32)
    idx[4100] = 1000 + 65535
    assert pack(idx)[0] == 1
    for pos in (0, 4095, 4096, 8999):
        bad = np.full(9000, 1000, dtype=np.int32)
        bad[pos] = 1000 + 65536 if pos else 1000 - 65536
        assert pack(bad)[0] == 0, pos
    assert pack(np.array([-5, 7, 65530 - 5], dtype=np.int32))[0] == 1   # negative (invalid) indices must not break the packer
