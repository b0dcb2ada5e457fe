"""The device BVH construction (device/dbvhbuild.cuh) executed on the CPU -- launches as loops, device
memory as host memory (tests/native/bvhbuild_on_host.cpp) -- against the host front end's recursive build, which restates
BVHAccel::recursiveBuild + flattenBVHTree with the reference's own libstdc++ std::partition / std::nth_element and is
itself pinned through the reference's images and its aggregatetest (test_oracle_pinning.py).  Same nodes, same depth-first
order, same primitive order in the leaves; floats compared by value (a zero's sign may differ, see dbvhbuild.cuh)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import pbrt_v2_b200 as pb
from helpers import host_scene

CUDA_INC = "/usr/local/cuda/include"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")), reason="CUDA headers not installed")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("bvhb") / "libbvhbuild_on_host.so")
    src = os.path.join(pb.REPO, "tests", "native", "bvhbuild_on_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden", "-Wl,-Bsymbolic",
                    "-I" + CUDA_INC, src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.doh_bvh_build.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p,
                                  C.POINTER(C.c_uint32), C.c_uint, C.c_int]
    return lib


def emu_build(emu, bounds, maxp, blocks=0, reverse=0):
    bounds = np.ascontiguousarray(bounds, np.float32)
    n = len(bounds)
    nodes = np.zeros(max(2 * n, 1), pb.NODE_DTYPE)
    order = np.zeros(max(n, 1), np.uint32)
    nn, nl = C.c_uint32(), C.c_uint32()
    assert emu.doh_bvh_build(bounds.ctypes.data, n, maxp, nodes.ctypes.data, C.byref(nn), order.ctypes.data, C.byref(nl), blocks,
                             reverse) == 0
    return nodes[:nn.value], order[:n], nl.value


def same_tree(a, b):
    return len(a) == len(b) and all(np.array_equal(a[k], b[k]) for k in ("bmin", "bmax", "offset", "nPrimitives", "axis"))


def boxes(kind, n, rs):
    f32 = np.float32
    if kind == "random":      # separated small boxes: every leaf holds one primitive
        c, e = rs.rand(n, 3).astype(f32) * 10 - 5, rs.rand(n, 3).astype(f32) ** 3
    elif kind == "huge":      # boxes far larger than their spread: the SAH prefers leaves (bvh.cpp:343-361)
        c, e = rs.rand(n, 3).astype(f32), np.full((n, 3), 100, f32)
    elif kind == "dup":       # coincident centroids: the degenerate branch (bvh.cpp:239-261)
        c, e = rs.randint(0, max(2, n // 8), (n, 1)).astype(f32) * np.ones((1, 3), f32), rs.rand(n, 3).astype(f32) * .1
    elif kind == "grid":      # many equal keys: ties in the partition predicate and in nth_element
        c, e = rs.randint(-4, 5, (n, 3)).astype(f32), np.full((n, 3), .5, f32)
    elif kind == "flat":
        c, e = rs.rand(n, 3).astype(f32), rs.rand(n, 3).astype(f32) * .05
        c[:, 2] = 0
        e[:, 2] = 0
    elif kind == "zeros":     # +0 / -0 coordinates
        c = np.zeros((n, 3), f32)
        c[rs.rand(n) < .5] *= -1
        e = rs.randint(0, 2, (n, 3)).astype(f32)
    elif kind == "line":
        c, e = np.zeros((n, 3), f32), np.full((n, 3), .25, f32)
        c[:, 0] = np.arange(n)
    elif kind == "skew":      # very uneven splits: deep trees
        c, e = (rs.rand(n, 3) ** 8).astype(f32) * 100, rs.rand(n, 3).astype(f32) * .01
    return np.concatenate([c - e, c + e], 1).astype(f32)


KINDS = ["random", "huge", "dup", "grid", "flat", "zeros", "line", "skew"]


@pytest.mark.parametrize("kind", KINDS)
def test_device_build_on_host_matches_recursive_build(emu, kind):
    rs = np.random.RandomState(100 + KINDS.index(kind))
    leaves = 0
    for n in [1, 2, 3, 4, 5, 6, 9, 33, 257, 3000]:
        for maxp in [1, 2, 4, 16, 255]:
            b = boxes(kind, n, rs)
            want_nodes, want_order = pb.host_bvh_build(b, maxp)
            # whole-grid schedule, a single block, three blocks run backwards: the atomics make no result depend on it
            for blocks, reverse in [(0, 0), (1, 0), (3, 1)]:
                nodes, order, _ = emu_build(emu, b, maxp, blocks, reverse)
                assert np.array_equal(order, want_order), (kind, n, maxp, blocks, reverse)
                assert same_tree(nodes, want_nodes), (kind, n, maxp, blocks, reverse)
            leaves = max(leaves, int(want_nodes["nPrimitives"].max()))
    if kind in ("huge", "dup", "grid", "zeros"):
        assert leaves > 1   # the multi-primitive leaf branches were really taken


def test_small_ranges_exhaustively(emu):
    """every key pattern of 2..5 primitives on a 3-value grid along x (ties included): libstdc++'s nth_element for <= 4
    elements (introselect round + insertion sort) and the first SAH partition above it, restated in the kernels"""
    import itertools
    f32 = np.float32
    checked = 0
    for n in (2, 3, 4, 5):
        for xs in itertools.product((0.0, 1.0, 2.0), repeat=n):
            if len(set(xs)) == 1:
                continue   # (degenerate centroid bounds: covered by "dup")
            c = np.zeros((n, 3), f32)
            c[:, 0] = xs
            c[:, 1] = np.arange(n) * 0.125   # distinct boxes, x stays the widest axis
            e = np.full((n, 3), 0.25, f32)
            b = np.concatenate([c - e, c + e], 1).astype(f32)
            for maxp in (1, 4):
                want_nodes, want_order = pb.host_bvh_build(b, maxp)
                nodes, order, _ = emu_build(emu, b, maxp)
                assert np.array_equal(order, want_order) and same_tree(nodes, want_nodes), (xs, maxp)
                checked += 1
    assert checked > 600


def test_random_campaign(emu):
    rs = np.random.RandomState(77)
    for it in range(400):
        n = int(rs.randint(1, 400))
        kind = KINDS[rs.randint(len(KINDS))]
        b = boxes(kind, n, rs)
        if rs.rand() < .3:   # mixtures: part of another family, shuffled in
            b2 = boxes(KINDS[rs.randint(len(KINDS))], n, rs)
            m = rs.rand(n) < .5
            b[m] = b2[m]
        maxp = int(rs.choice([1, 2, 3, 4, 8, 255]))
        want_nodes, want_order = pb.host_bvh_build(b, maxp)
        nodes, order, _ = emu_build(emu, b, maxp, int(rs.choice([0, 1, 2])), int(rs.randint(2)))
        assert np.array_equal(order, want_order) and same_tree(nodes, want_nodes), (it, kind, n, maxp)


def test_tree_invariants_and_levels(emu):
    rs = np.random.RandomState(5)
    n = 300000   # (several levels of scan partials)
    b = boxes("random", n, rs)
    nodes, order, levels = emu_build(emu, b, 4)
    assert np.array_equal(np.sort(order), np.arange(n))
    leaf = nodes["nPrimitives"] > 0
    assert nodes["nPrimitives"][leaf].sum() == n
    assert np.array_equal(np.sort(nodes["offset"][leaf]), np.cumsum(np.r_[0, nodes["nPrimitives"][leaf][np.argsort(nodes["offset"][leaf])]])[:-1])
    assert np.all(nodes["offset"][~leaf] > np.nonzero(~leaf)[0] + 1) and np.all(nodes["offset"][~leaf] < len(nodes))
    assert 12 < levels < 64   # BVHAccel::Intersect's todo[64] stack (bvh.cpp:411)
    want_nodes, want_order = pb.host_bvh_build(b, 4)
    assert np.array_equal(order, want_order) and same_tree(nodes, want_nodes)
    # root bounds = union of everything
    assert np.array_equal(nodes["bmin"][0], b[:, :3].min(0)) and np.array_equal(nodes["bmax"][0], b[:, 3:].max(0))


@pytest.mark.parametrize("base", ["bunny", "killeroo-simple"])
def test_config_scene_meshes(staged, emu, base):
    """the BASELINE scenes: the input the scene's own accelerator was built from -> the scene's own node array"""
    pb.keep_build_input(True)
    try:
        hs = host_scene(base, 32, 32, 4, 32, tag="-buildinput")   # (its own cache entry: loaded with the build input kept)
    finally:
        pb.keep_build_input(False)
    b = hs.build_input()
    assert len(b) == hs.counts()["primitives"] > 30000
    nodes, order, levels = emu_build(emu, b, 4)
    assert same_tree(nodes, hs.nodes())
    want_nodes, want_order = pb.host_bvh_build(b, 4)
    assert np.array_equal(order, want_order) and same_tree(nodes, want_nodes)
    nodes1, order1, _ = emu_build(emu, b, 1)   # maxPrimsInNode 1: what TransformedPrimitive's nested BVHAccel uses
    want_nodes, want_order = pb.host_bvh_build(b, 1)
    assert np.array_equal(order1, want_order) and same_tree(nodes1, want_nodes)


def test_entry_point_validates_before_touching_the_device():
    """argument checks of pbrtgpu_bvh_build run before the device is needed (and there is no CPU fallback behind them)"""
    b = boxes("random", 10, np.random.RandomState(1))
    for bad, kw in [(b, dict(split_method=pb.SPLIT_MIDDLE)), (b, dict(split_method=pb.SPLIT_EQUAL_COUNTS)),
                    (b, dict(max_prims_in_node=0)), (b, dict(max_prims_in_node=256))]:
        with pytest.raises(pb.PbrtGpuError, match="code [12]"):
            pb.bvh_build(bad, **kw)
    for v in (np.nan, np.inf, -np.inf):
        c = b.copy()
        c[3, 4] = v
        with pytest.raises(pb.PbrtGpuError, match="not finite"):
            pb.bvh_build(c)
    if pb.device_count() == 0:
        with pytest.raises(pb.PbrtGpuError, match="no CPU fallback"):
            pb.bvh_build(b)


def test_front_end_builder_option_has_no_silent_fallback(staged, monkeypatch):
    """Accelerator "bvh" "string builder" "gpu" (or PBRTGPU_BVH_BUILDER=gpu) must build on the device or fail: on a box
    without a GPU the scene does not load, and the message says why"""
    if pb.device_count() > 0:
        pytest.skip("a CUDA device is present: covered by test_zz_gpu_bvhbuild.py")
    monkeypatch.setenv("PBRTGPU_BVH_BUILDER", "gpu")
    with pytest.raises(pb.PbrtGpuError, match="no CUDA device"):
        host_scene("bunny", 24, 24, 4, 2, tag="-devbuilt-cpu")
    monkeypatch.setenv("PBRTGPU_BVH_BUILDER", "host")
    assert host_scene("bunny", 24, 24, 4, 2, tag="-hostbuilt").counts()["nodes"] > 0
