"""CPU: the product's device math (the __host__ __device__ functions of include/swr/detail/*.cuh)
executed on the host by tests/hostcheck/hostcheck.cu and diffed against the oracle.  This proves
the arithmetic of the CUDA path without a GPU; the -m gpu tests prove the orchestration."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import common
from softwarerenderer_b200 import scenes as S

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cu")
SO = os.path.join(HERE, "hostcheck", "libhostcheck.so")
SUPPORTED_PS = (S.PS_COUNT_ID, S.PS_GOURAUD, S.PS_GOURAUD_DEPTH, S.PS_VARY_DUMP, S.PS_TEXTURED_ANISO)


@pytest.fixture(scope="module")
def hostcheck(oracle):
    deps = [SRC, os.path.join(ROOT, "include", "swr", "Texture.h")] + [os.path.join(ROOT, "include", "swr", "detail", f) for f in ("common.h", "geometry.cuh", "tile.cuh")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.run(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17",
                        "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-I" + os.path.join(ROOT, "include"),
                        "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.hostcheck_draw.argtypes = [C.POINTER(oracle.SwrScene)]
    lib.hostcheck_draw.restype = C.c_int

    def run(scene):
        targets = oracle.fresh_targets(scene.width, scene.height)
        s, keep, _ = oracle._fill(scene, targets, 0)
        assert lib.hostcheck_draw(C.byref(s)) == 0
        out = dict(targets)
        out["fragments"] = int(s.fragments)
        return out
    return run


SCENES = [(l, s) for l, s in common.parity_scenes(ntri=512) if s.ps in SUPPORTED_PS]


@pytest.mark.parametrize("label,scene", SCENES, ids=[l for l, _ in SCENES])
def test_device_math_on_host_matches_oracle(oracle, hostcheck, label, scene):
    got, want = hostcheck(scene), oracle.run(scene, "oracle")
    assert got["fragments"] == want["fragments"]
    assert not common.diff_buffers(got, want), label


FUZZ = [  # (class of triangles, how many, surface)
    (0, 150000, 3840, 2160), (2, 100000, 3840, 2160), (4, 100000, 3840, 2160), (5, 100000, 3840, 2160), (6, 100000, 3840, 2160),
    (1, 300, 3840, 2160), (3, 5000, 3840, 2160), (7, 300, 3840, 2160),
    (0, 100000, 16384, 16384), (2, 50000, 16384, 16384), (3, 3000, 16384, 16384), (4, 50000, 16384, 16384), (6, 50000, 16384, 16384),
]


@pytest.mark.parametrize("kind,n,w,h", FUZZ, ids=[f"kind{k}_{w}" for k, _, w, _ in FUZZ])
def test_certified_pixel_bounds_never_drop_a_fragment(hostcheck, kind, n, w, h):
    """geometry.cuh tightPixelBounds: over random tiny / block-sized / screen-sized / sliver / pixel-centre /
    pixel-corner / sub-pixel / guard-band triangles, no pixel the reference's full block walk covers lies outside the
    certified bounds, and the rectangle-limited coverage equals the full one (0 violations)."""
    lib = C.CDLL(SO)
    f = lib.hostcheck_tight_box_fuzz
    f.argtypes = [C.c_int, C.c_long, C.c_ulonglong, C.c_int, C.c_int, C.POINTER(C.c_long), C.POINTER(C.c_long)]
    f.restype = C.c_long
    checked, excluded = C.c_long(), C.c_long()
    bad = f(kind, n, 1000 + kind, w, h, C.byref(checked), C.byref(excluded))
    assert bad == 0
    assert checked.value > n // 2 and excluded.value > 0


def test_device_math_without_the_certified_bounds(oracle, hostcheck, monkeypatch):
    """The same records with the reference's 8-aligned block boxes (GeomArgs::noTightBox) give the same pixels."""
    monkeypatch.setenv("HOSTCHECK_NO_TIGHT_BOX", "1")
    for label, scene in SCENES[::9]:
        got, want = hostcheck(scene), oracle.run(scene, "oracle")
        assert got["fragments"] == want["fragments"] and not common.diff_buffers(got, want), label
