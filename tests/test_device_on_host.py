"""The product's device headers (dscene.cuh, dshade.cuh, danim.cuh, dmath.cuh, refmath.h) compiled for the HOST
(tests/native/device_on_host.cpp: g++ with the CUDA intrinsics shimmed, no FMA contraction) and compared with the oracle
bit for bit: BVH walks in both forms (Traverse<>, the wavefront's TravStep<>), TransformedPrimitives, hit geometry, and
the shading steps of the path state machine (shading frame, light sampling, BSDF::f / Pdf / Sample_f incl. the measured
BRDF's k-nearest + gather form).  No GPU needed: this is the CPU-stage guard for edits to the device arithmetic."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import pbrt_v2_b200 as pb
import oracle_api as oa
from helpers import host_scene, random_rays, scene_bounds

CUDA_INC = "/usr/local/cuda/include"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")), reason="CUDA headers not installed")


@pytest.fixture(scope="module")
def doh(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("doh") / "libdevice_on_host.so")
    src = os.path.join(pb.REPO, "tests", "native", "device_on_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I" + CUDA_INC, src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.doh_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.doh_intersect_p.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.doh_shade_test.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return lib


def _rays(hs, base, n, seed):
    lo, hi = scene_bounds(hs)
    if base.startswith("anim"):
        lo, hi = np.maximum(lo, [-300, -300, -150]).astype(np.float32), np.minimum(hi, [300, 300, 150]).astype(np.float32)
    rays = random_rays(n, seed, lo, hi)
    rays["time"] = np.random.RandomState(seed + 1).choice(np.array([-0.5, 0.0, 0.25, 0.5, 0.999, 1.0, 1.5], np.float32), n)
    rays["time"][::3] = np.random.RandomState(seed + 2).rand(len(rays[::3])).astype(np.float32)
    return rays


@pytest.mark.parametrize("base", ["bunny", "killeroo-simple", "anim-killeroos-moving"])
@pytest.mark.parametrize("stepwise", [0, 1])
def test_device_traversal_on_host_matches_oracle(staged, doh, base, stepwise):
    hs = host_scene(base, 32, 32, 4, 32)
    rays = _rays(hs, base, 100000, 7)
    want, occ_want = oa.intersect(hs, rays), oa.intersect_p(hs, rays)
    got = np.zeros(len(rays), dtype=pb.HIT_DTYPE)
    occ = np.zeros(len(rays), np.uint8)
    doh.doh_intersect(hs.desc_ptr, rays.ctypes.data, len(rays), got.ctypes.data, stepwise)
    doh.doh_intersect_p(hs.desc_ptr, rays.ctypes.data, len(rays), occ.ctypes.data, stepwise)
    assert np.array_equal(got["hit"], want["hit"]) and np.array_equal(occ, occ_want)
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    h = want["hit"] == 1
    assert h.mean() > 0.2
    for f in ("primitive", "b1", "b2", "p", "n"):
        a, b = got[f][h], want[f][h]
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f


@pytest.mark.parametrize("base", ["bunny", "killeroo-simple", "anim-killeroos-moving"])
def test_device_shading_on_host_matches_oracle(staged, doh, base):
    hs = host_scene(base, 32, 32, 4, 32)
    n = 60000
    rays = _rays(hs, base, n, 9)
    u = np.random.RandomState(4).rand(n, 6).astype(np.float32)
    a, b = np.zeros((n, 36), np.float32), np.zeros((n, 36), np.float32)
    oa.lib().oracle_shade_test.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    oa.lib().oracle_shade_test(hs.desc_ptr, rays.ctypes.data, n, u.ctypes.data, a.ctypes.data)
    doh.doh_shade_test(hs.desc_ptr, rays.ctypes.data, n, u.ctypes.data, b.ctypes.data)
    assert (a[:, 15] != 0).mean() > 0.2 and (a[:, 34] != 0).mean() > 0.1   # hits shaded, BSDF samples drawn
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), "columns differing: %s" % np.nonzero(~same.all(axis=0))[0].tolist()
