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
