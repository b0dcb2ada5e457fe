"""The LD sample-table kernels (device/dtables.cuh) executed on the CPU under a warp emulation
(tests/native/tables_on_host.cpp: 32 host threads = the lanes of one warp, __syncwarp() = a barrier) and compared with the
oracle's LDPixelSample bit for bit.  Covers what the GPU-less stage otherwise cannot: every samples-per-pixel count the ABI
accepts -- including the ones whose 26 index arrays no longer fit shared memory at once and are shuffled in groups (2048 spp
and up; BASELINE C3 renders at 4096) -- both kernels, general Sample layouts, and the RNG hand-over from pixel to pixel."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import pbrt_v2_b200 as pb
import oracle_api as oa
from helpers import layout_rd

CUDA_INC = "/usr/local/cuda/include"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")), reason="CUDA headers not installed")


def _build(tmp_path_factory, name, defines=()):
    so = str(tmp_path_factory.mktemp(name) / ("lib%s.so" % name))
    src = os.path.join(pb.REPO, "tests", "native", "tables_on_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden", "-Wl,-Bsymbolic",
                    "-pthread", "-I" + CUDA_INC] + list(defines) + [src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.toh_ld_samples.argtypes = [C.c_void_p, C.c_int] + [C.c_int] * 4 + [C.c_int, C.c_int, C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def toh(tmp_path_factory):
    return _build(tmp_path_factory, "tables_on_host")


def _path_rd(spp, res=64, ntasks=16):
    rd = pb.RenderDesc()
    rd.xResolution = rd.yResolution = res
    rd.cropWindow[:] = [0, 1, 0, 1]
    rd.filterXWidth = rd.filterYWidth = 0.5
    rd.samplesPerPixel = spp
    rd.maxDepth = 5
    rd.nTasks, rd.taskBegin, rd.taskEnd = ntasks, 0, ntasks
    rd.shutterOpen, rd.shutterClose = 0.25, 0.75
    pb.set_path_layout(rd)
    return rd


def _run(toh, rd, task, general=0, rounds=0):
    win = oa.task_window(rd, task)   # x0, x1, y0, y1
    out = np.zeros((rd.samplesPerPixel, pb.sample_floats(rd)), np.float32)
    rc = toh.toh_ld_samples(C.addressof(rd), task, win[0], win[2], win[1], win[3], general, rounds, out.ctypes.data)
    assert rc == 0
    return out


def _same(a, b):
    return np.array_equal(a.view(np.uint32), b.view(np.uint32))


@pytest.mark.parametrize("spp", [1, 2, 4, 32, 256, 1024, 2048, 4096, 8192])
def test_path_table_kernel_matches_oracle(toh, spp):
    rd = _path_rd(spp)
    for task in ([0, 5] if spp <= 1024 else [3]):
        assert _same(_run(toh, rd, task), oa.ld_samples(rd, task)), (spp, task)


@pytest.mark.parametrize("spp", [1, 2, 4, 8, 16, 32, 256, 1024, 2048, 4096, 8192])
def test_path_table_kernel_in_two_launches_matches_oracle(toh, spp):
    """the render path's form: the draw-stream walk (stages = 1) and the shuffles + rows (stages = 2) as separate launches,
    the latter with its own shared-memory plan; also across pixels (the RNG state handed on through rngCur / rngNext / mti)"""
    toh.toh_set_split(1)
    try:
        rd = _path_rd(spp)
        for task in ([0, 5] if spp <= 1024 else [3]):
            assert _same(_run(toh, rd, task), oa.ld_samples(rd, task)), (spp, task)
        if spp in (4, 32, 2048):
            assert _same(_run(toh, rd, 2, rounds=3), oa.ld_samples(rd, 2, 3))
    finally:
        toh.toh_set_split(0)


@pytest.mark.parametrize("spp", [1, 8, 128, 2048])
def test_general_table_kernel_on_path_layout_matches_oracle(toh, spp):
    rd = _path_rd(spp)
    assert _same(_run(toh, rd, 7, general=1), oa.ld_samples(rd, 7))


def test_general_table_kernel_matches_golden_layouts(toh):
    g = np.load(os.path.join(pb.REPO, "tests", "golden", "ld_samples_general.npz"))
    names = sorted(k[:-5] for k in g.files if k.endswith("_meta"))
    assert names
    for name in names:
        rd, task = layout_rd(g[name + "_meta"])
        got = _run(toh, rd, task)
        assert _same(got, oa.ld_samples(rd, task)), name
        assert _same(got, g[name].reshape(got.shape)), name   # the reference's own LDPixelSample output


@pytest.mark.parametrize("general", [0, 1])
@pytest.mark.parametrize("spp,pixel", [(4, 9), (64, 5), (2048, 1)])
def test_rng_stream_hands_over_from_pixel_to_pixel(toh, spp, pixel, general):
    rd = _path_rd(spp)
    assert _same(_run(toh, rd, 2, general=general, rounds=pixel), oa.ld_samples(rd, 2, pixel))


def test_large_arrays_and_spp_in_general_kernel(toh):
    rd = _path_rd(2048)
    rd.integrator = 1
    rd.n1D, rd.n2D = 2, 3
    rd.size1D[0], rd.size1D[1] = 1, 32
    rd.size2D[0], rd.size2D[1], rd.size2D[2] = 16, 1, 32
    assert _same(_run(toh, rd, 1), oa.ld_samples(rd, 1))   # 32 x 2048 = 65536 table entries: the 16-bit index limit
    rd.samplesPerPixel = 256
    assert _same(_run(toh, rd, 1, rounds=2), oa.ld_samples(rd, 1, 2))


def test_paired_shuffle_steps_give_the_same_tables(tmp_path_factory):
    """-DPBRT_TABLES_PAIRED=1 (a round-2 kernel experiment, not in the shipped kernels): k_tables runs its Fisher-Yates
    chains two steps per shared-memory round trip, patching in registers what the second step would have read after the
    first one's stores.  Same tables, bit for bit, incl. spp 1 (no pair) and the grouped path of 2048 spp."""
    lib = _build(tmp_path_factory, "tables_on_host_paired", ["-DPBRT_TABLES_PAIRED=1"])
    assert lib.toh_paired_build() == 1
    for spp, tasks, rounds in [(1, [0, 3], 2), (2, [1], 3), (4, [2], 1), (8, [5], 0), (32, [0, 7], 1), (256, [4], 0), (1024, [6], 0),
                               (2048, [3], 0)]:
        rd = _path_rd(spp)
        for task in tasks:
            assert _same(_run(lib, rd, task, rounds=rounds), oa.ld_samples(rd, task, rounds)), (spp, task)
