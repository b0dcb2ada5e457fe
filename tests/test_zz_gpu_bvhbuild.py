"""pbrtgpu_bvh_build on the GPU against the host front end's recursive build: same nodes, same
depth-first order, same primitive order; and a scene whose accelerator was built on the device renders the same film.
(Named to sort last: the kernels were first verified on the CPU -- test_bvhbuild_on_host.py -- and these are their
first runs on hardware.)"""
import os

import numpy as np
import pytest

import pbrt_v2_b200 as pb
from helpers import host_scene
from test_bvhbuild_on_host import KINDS, boxes, same_tree

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", KINDS)
def test_device_build_matches_recursive_build(kind):
    rs = np.random.RandomState(200 + KINDS.index(kind))
    for n in [1, 2, 5, 33, 257, 3000, 100000]:
        for maxp in [1, 4, 255]:
            b = boxes(kind, n, rs)
            want_nodes, want_order = pb.host_bvh_build(b, maxp)
            nodes, order, st = pb.bvh_build(b, maxp)
            assert np.array_equal(order, want_order), (kind, n, maxp)
            assert same_tree(nodes, want_nodes), (kind, n, maxp)
            assert st.nLevels >= 1


def test_repeated_builds_are_identical():
    b = boxes("grid", 50000, np.random.RandomState(3))
    first = pb.bvh_build(b, 4)
    for _ in range(3):
        again = pb.bvh_build(b, 4)
        assert again[0].tobytes() == first[0].tobytes() and np.array_equal(again[1], first[1])


@pytest.mark.parametrize("base", ["bunny", "killeroo-simple"])
def test_config_scene_built_on_device_is_the_same_scene(staged, base, monkeypatch):
    host = host_scene(base, 48, 48, 4, 8)
    monkeypatch.setenv("PBRTGPU_BVH_BUILDER", "gpu")
    dev = host_scene(base, 48, 48, 4, 8, tag="-devbuilt")   # (the tag keeps helpers' scene cache from returning `host`)
    monkeypatch.delenv("PBRTGPU_BVH_BUILDER")
    assert dev is not host
    assert same_tree(dev.nodes(), host.nodes())
    px_h, _ = host.render()
    px_d, _ = dev.render()
    assert px_h.tobytes() == px_d.tobytes()


def test_large_build_and_timing():
    """2M random boxes (a C5-like input): same tree, and the device build's time is reported next to the host's"""
    import time
    n = 2000000
    b = boxes("random", n, np.random.RandomState(11))
    t0 = time.time()
    want_nodes, want_order = pb.host_bvh_build(b, 4)
    host_s = time.time() - t0
    pb.bvh_build(b[:1000], 4)   # context / module load outside the timing
    nodes, order, st = pb.bvh_build(b, 4)
    assert np.array_equal(order, want_order) and same_tree(nodes, want_nodes)
    print("\nBVH build, %d primitives: host %.2f s, device %.1f ms (%d levels)" % (n, host_s, st.buildMs, st.nLevels))
