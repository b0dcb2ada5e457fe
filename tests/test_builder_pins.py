"""Direct pins of the two scene-preparation builders on the reference's OWN arrays (missing #8): the
accelerator's `LinearBVHNode[]` + reordered primitive list (accelerators/bvh.cpp:153-202, 383-399) and the TriangleMesh that
`LoopSubdiv::Refine` creates (shapes/loopsubdiv.cpp:221-422), dumped by oracle/_ref/bin/refdump (`bvh`, `loopsubdiv` modes:
our tool linked against the unmodified reference library) -- committed as tests/golden/builders.npz
(tests/golden/make_builder_golden.py), and taken live on the config scenes' inputs where the reference is built.

Pinned here: the host front end's builders (host/bvh.cpp, host/shapes.cpp).  The device builders are compared with those,
input for input, by test_bvhbuild_on_host.py / test_loopsubdiv_on_host.py (CPU) and tests/test_zz_gpu_*.py (GPU), and one
case each is compared with the reference's dump directly below."""
import os
import sys

import numpy as np
import pytest

import pbrt_v2_b200 as pb

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_builder_golden as mg   # noqa: E402
import make_envmap_golden as me   # noqa: E402

GOLDEN = np.load(os.path.join(HERE, "golden", "builders.npz"))
live = pytest.mark.skipif(not os.path.exists(mg.REFDUMP), reason="the reference is not built here (oracle/_ref)")


def same_tree(ours, our_order, ref, ref_order):
    """Our NODE_DTYPE records + primitive order against the reference's LinearBVHNode dump + reordered primitive list: every
    node's bounds, interior nodes' axis and second-child offset, leaves' primitive count and CONTENT (the input
    indices of its primitives, in order).

    What is NOT compared is where a leaf's block lies in the reordered primitive array: the reference builds both children
    inside one call's argument list -- `node->InitInterior(dim, recursiveBuild(.. start, mid ..), recursiveBuild(.. mid, end ..))`,
    bvh.cpp:257-261, 367-373 -- and C++ leaves the order of the two unspecified: the g++ build this pin runs against builds
    the RIGHT child first (the first leaf in depth-first order gets the largest primitivesOffset), a left-to-right compiler
    gives the order the front end uses.  The walk visits the same primitives in the same order either way."""
    if len(ours) != len(ref) or len(our_order) != len(ref_order):
        return False
    # (bounds by value: where a box face lies at zero the builders may keep -0 or +0 -- std::min / std::max return their first
    # argument on a tie and the operand order of a Union is not part of what is pinned; bounds only enter comparisons)
    if not (np.array_equal(ours["bmin"], ref["bmin"]) and np.array_equal(ours["bmax"], ref["bmax"]) and
            np.array_equal(ours["nPrimitives"], ref["nPrimitives"])):
        return False
    interior = ref["nPrimitives"] == 0
    if not (np.array_equal(ours["axis"][interior], ref["axis"][interior]) and np.array_equal(ours["offset"][interior], ref["offset"][interior])):
        return False
    for i in np.nonzero(~interior)[0]:
        n = int(ref["nPrimitives"][i])
        a, b = int(ours["offset"][i]), int(ref["offset"][i])
        if not np.array_equal(our_order[a:a + n], ref_order[b:b + n]):
            return False
    # (and with g++'s right-first evaluation the reference's blocks are ours mirrored: leaf k from the front = leaf k from the back)
    return True


def same_mesh(ours, ref):
    return (np.array_equal(ours[2], ref[2]) and ours[0].shape == ref[0].shape and np.array_equal(ours[0], ref[0], equal_nan=True) and
            np.array_equal(ours[1], ref[1], equal_nan=True))


@pytest.fixture(params=[1 << 30, 2], ids=["serial", "forked"])
def build_mode(request):
    """the host build as one recursion, and with its subtrees forked onto threads from the root down (host/bvh.cpp
    ParallelBuild; production builds fork above 131 072 primitives): the same tree either way"""
    pb.set_bvh_parallel_min(request.param)
    yield request.param
    pb.set_bvh_parallel_min()


def test_host_bvh_build_is_the_references_tree_golden(build_mode):
    keys = sorted(k[:-7] for k in GOLDEN.files if k.startswith("bvh_") and k.endswith("_bounds"))
    assert len(keys) >= 32
    multi = 0
    for k in keys:
        maxp = int(k.rsplit("_", 1)[1])
        nodes, order = pb.host_bvh_build(GOLDEN[k + "_bounds"], maxp)
        assert same_tree(nodes, order, GOLDEN[k + "_nodes"], GOLDEN[k + "_order"]), k
        multi = max(multi, int(nodes["nPrimitives"].max()))
    assert multi > 1   # multi-primitive leaves are among the pinned trees


@pytest.mark.parametrize("split", [pb.SPLIT_SAH, 0, 1])
def test_forked_host_build_is_the_serial_build(split):
    """every split method ("sah", "middle" = 0, "equal" = 1), random and degenerate inputs, leaf sizes 1 / 4 / 255: node array and
    primitive order byte for byte"""
    from test_bvhbuild_on_host import boxes, KINDS
    rs = np.random.RandomState(11)
    for kind in KINDS:
        for n, maxp in [(2, 1), (37, 4), (1000, 1), (5000, 4), (3000, 255)]:
            b = boxes(kind, n, rs)
            pb.set_bvh_parallel_min(1 << 30)
            n0, o0 = pb.host_bvh_build(b, maxp, split)
            pb.set_bvh_parallel_min(2)
            try:
                n1, o1 = pb.host_bvh_build(b, maxp, split)
            finally:
                pb.set_bvh_parallel_min()
            assert np.array_equal(n0.view(np.uint8), n1.view(np.uint8)) and np.array_equal(o0, o1), (kind, n, maxp)


def test_host_loopsubdiv_refine_is_the_references_mesh_golden():
    keys = sorted(k[:-4] for k in GOLDEN.files if k.startswith("subdiv_") and k.endswith("_idx"))
    assert len(keys) >= 30
    for k in keys:
        levels = int(k.rsplit("_L", 1)[1])
        cm = pb.ControlMesh(GOLDEN[k + "_idx"], GOLDEN[k + "_P"])
        assert same_mesh(cm.host_refine(levels), (GOLDEN[k + "_rP"], GOLDEN[k + "_rN"], GOLDEN[k + "_ridx"])), k


@live
@pytest.mark.parametrize("base", ["bunny", "killeroo-simple"])
def test_scene_accelerator_is_the_references_tree_live(staged, base, build_mode):
    """the config scenes' own build input (69 k / 66 k refined primitives in the order BVHAccel gets them; the killeroo's are
    the refined loopsubdiv mesh): the reference's BVHAccel over those bounds IS the accelerator the front end built for the
    scene (maxPrimsInNode 4), and with maxPrimsInNode 1 (what nested TransformedPrimitive BVHs use) the builders agree too"""
    from helpers import host_scene
    pb.keep_build_input(True)
    try:
        hs = host_scene(base, 16, 16, 4, 1, tag="-pin")
    finally:
        pb.keep_build_input(False)
    b = hs.build_input()
    assert len(b) > 60000
    for maxp in (4, 1):
        ref_nodes, ref_order = mg.ref_bvh(b, maxp)
        nodes, order = pb.host_bvh_build(b, maxp)
        assert same_tree(nodes, order, ref_nodes, ref_order), (base, maxp)
        if maxp == 4:   # ... and that IS the accelerator the front end built for the scene
            sn = hs.nodes()
            assert all(np.array_equal(sn[f], nodes[f]) for f in ("bmin", "bmax", "offset", "nPrimitives", "axis")), base


@live
def test_host_loopsubdiv_refine_is_the_references_mesh_live(staged):
    """the killeroo's control mesh at the shipped nlevels (1), at configs[3]'s (3) and in between"""
    from test_loopsubdiv_on_host import _killeroo_control_mesh
    idx, P = _killeroo_control_mesh()
    cm = pb.ControlMesh(idx, P)
    for levels in (1, 2, 3):
        assert same_mesh(cm.host_refine(levels), mg.ref_loopsubdiv(idx, P, levels)), levels


# ---- InfiniteAreaLight's constructor (lights/infinite.cpp:67-104): host/envmap.cpp against the reference's own members ----
def _same_envmap(mine, head, data):
    ref = pb.parse_envmap_dump(head, data)
    for k in ("width", "height", "nu", "nv"):
        assert mine[k] == ref[k], k
    for k in ("texels", "func", "cdf", "funcInt", "margCdf", "margFuncInt"):
        assert np.array_equal(mine[k].view(np.uint32), ref[k].view(np.uint32)), k


def test_host_envmap_build_is_the_references_light_golden(tmp_path):
    """radianceMap level 0 (ReadImagePFM -> x L -> MIPMap: Lanczos resampling for the 12x7 and 5x3 files) and the
    Distribution2D over the filtered luminance x sin(theta), bit for bit: RGB / grey files, both byte orders, a scale word"""
    g = np.load(os.path.join(HERE, "golden", "envmaps.npz"))
    names = sorted(k for k in g.files if k.endswith("_pfm") or k.endswith("_tga"))
    assert len(names) >= 12
    for key in names:   # (Targa files: ReadImageTGA over targa.c -- raw / RLE, grey, BGR(A), colour-mapped, all scan directions)
        k, ext = key[:-4], key[-3:]
        path = str(tmp_path / (k + "." + ext))
        open(path, "wb").write(g[key].tobytes())
        _same_envmap(pb.host_envmap_build(path, g[k + "_L"]), g[k + "_head"], g[k + "_data"])
    _same_envmap(pb.host_envmap_build(None, g["nomap_L"]), g["nomap_head"], g["nomap_data"])


@live
def test_host_envmap_build_is_the_references_light_live(tmp_path):
    """random resolutions up to 70 x 40 (one- and two-sided resampling, single rows and columns)"""
    from helpers import write_env_pfm
    rs = np.random.RandomState(7)
    sizes = [(1, 2), (2, 1), (3, 1), (1, 5), (7, 1), (64, 32), (17, 9), (70, 40), (33, 8), (8, 33)]
    sizes += [(int(rs.randint(1, 40)), int(rs.randint(1, 24))) for _ in range(14)]
    for i, (w, h) in enumerate(sizes):
        path = str(tmp_path / ("m%d.pfm" % i))
        write_env_pfm(path, w, h, 100 + i, 3 if i % 3 else 1, [-1.0, 1.0, -2.5][i % 3])
        L = rs.uniform(0, 2, 3).astype(np.float32)
        _same_envmap(pb.host_envmap_build(path, L), *me.ref_envmap(path, L))
    from helpers import write_tga
    for i in range(24):   # Targa: every kind x raw / RLE x the four scan directions
        kind = ["bgr", "mono", "mapped"][i % 3]
        path = str(tmp_path / ("t%d.tga" % i))
        write_tga(path, int(rs.randint(1, 20)), int(rs.randint(1, 12)), 200 + i, kind, bool(i & 1), bool(i & 2), bool(i & 4),
                  bool(i & 8) and kind != "mono", int(rs.randint(0, 5)))
        L = rs.uniform(0, 2, 3).astype(np.float32)
        _same_envmap(pb.host_envmap_build(path, L), *me.ref_envmap(path, L))
