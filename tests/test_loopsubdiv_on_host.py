"""LoopSubdiv::Refine's device version (device/dloopsubdiv.cuh) executed on the CPU -- launches as loops,
device memory as host memory (tests/native/loopsubdiv_on_host.cpp) -- against the host front end's LoopSubdiv::Refine, the
restatement of shapes/loopsubdiv.cpp that the oracle-vs-reference image tests and the random-scene campaign pin (the
killeroo of the BASELINE scenes is a loopsubdiv shape).  Same vertices in the same order, same limit positions and normals
bit for bit, same triangles."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import pbrt_v2_b200 as pb

CUDA_INC = "/usr/local/cuda/include"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")), reason="CUDA headers not installed")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("lsd") / "libloopsubdiv_on_host.so")
    src = os.path.join(pb.REPO, "tests", "native", "loopsubdiv_on_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden", "-Wl,-Bsymbolic",
                    "-I" + CUDA_INC, src, "-o", so], check=True)
    lib = C.CDLL(so)
    lib.doh_loopsubdiv_refine.argtypes = [C.POINTER(pb.SubdivMesh), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint, C.c_int]
    return lib


def emu_refine(emu, cm, nlevels, blocks=0, reverse=0):
    nv, nt = cm.counts(nlevels)
    P, N, idx = np.zeros((nv, 3), np.float32), np.zeros((nv, 3), np.float32), np.zeros((nt, 3), np.int32)
    assert emu.doh_loopsubdiv_refine(C.byref(cm.desc), nlevels, P.ctypes.data, N.ctypes.data, idx.ctypes.data, blocks, reverse) == 0
    return P, N, idx


def same_mesh(a, b):
    """indices equal; positions and normals equal as floats, bit for bit up to the sign of a zero (the host hook returns P
    through TriangleMesh's identity object-to-world transform, which turns -0 into +0)"""
    return (np.array_equal(a[2], b[2]) and a[0].shape == b[0].shape and np.array_equal(a[0], b[0], equal_nan=True) and
            np.array_equal(a[1], b[1], equal_nan=True))


def fan(n, closed, rs, centre_z=0.5):
    """n triangles around vertex 0: interior vertex of valence n (closed) or boundary vertex of valence n + 1 (open)"""
    ang = np.linspace(0, 2 * np.pi * (1 if closed else 0.8), n + (0 if closed else 1), endpoint=not closed)
    ring = np.stack([np.cos(ang), np.sin(ang), 0.1 * rs.rand(len(ang))], 1)
    P = np.vstack([[0, 0, centre_z], ring]).astype(np.float32)
    m = len(ring)
    tris = [[0, 1 + i, 1 + (i + 1) % m] for i in range(n if closed else m - 1)]
    return tris, P


def grid(nx, ny, rs, holes=0.0):
    xs, ys = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1))
    P = np.stack([xs.ravel(), ys.ravel(), rs.rand(xs.size) * 0.7], 1).astype(np.float32)
    tris = []
    for y in range(ny):
        for x in range(nx):
            a, b, c, d = y * (nx + 1) + x, y * (nx + 1) + x + 1, (y + 1) * (nx + 1) + x + 1, (y + 1) * (nx + 1) + x
            if rs.rand() < holes:
                continue
            tris += [[a, b, c], [a, c, d]] if rs.rand() < .5 else [[a, b, d], [b, c, d]]
    return tris, P


def closed_solid(kind, rs):
    if kind == "tet":
        return [[0, 1, 2], [0, 3, 1], [0, 2, 3], [1, 3, 2]], np.float32([[1, 1, 1], [-1, -1, 1], [-1, 1, -1], [1, -1, -1]])
    # octahedron
    P = np.float32([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]]) + (rs.rand(6, 3) * .2).astype(np.float32)
    return [[0, 2, 4], [2, 1, 4], [1, 3, 4], [3, 0, 4], [2, 0, 5], [1, 2, 5], [3, 1, 5], [0, 3, 5]], P


def control_meshes():
    rs = np.random.RandomState(42)
    out = [("tet", closed_solid("tet", rs)), ("octahedron", closed_solid("oct", rs))]
    for n in (3, 4, 5, 6, 7, 12, 33):   # interior valences: 3 (its own beta), regular 6, extraordinary ones, a large one
        out.append(("fan-closed-%d" % n, fan(n, True, rs)))
    for n in (1, 2, 3, 4, 5, 9):        # boundary valences 2, 3, 4 (regular), 5, 6, 10: every branch of the boundary tangent
        out.append(("fan-open-%d" % n, fan(n, False, rs)))
    out.append(("grid", grid(6, 5, rs)))
    out.append(("grid-holes", grid(9, 8, rs, holes=0.25)))   # boundaries inside, bow-tie vertices, isolated quads
    tris, P = grid(4, 4, rs)
    out.append(("unused-vertices", (tris, np.vstack([P, rs.rand(3, 3).astype(np.float32)]))))
    return out


@pytest.mark.parametrize("name,mesh", control_meshes(), ids=[n for n, _ in control_meshes()])
def test_device_refine_on_host_matches_host_refine(emu, name, mesh):
    cm = pb.ControlMesh(*mesh)
    for nlevels in (0, 1, 2, 3):
        want = cm.host_refine(nlevels)
        assert (len(want[0]), len(want[2])) == cm.counts(nlevels), (name, nlevels)
        # whole-grid schedule, a single block, three blocks run backwards
        for blocks, reverse in [(0, 0), (1, 0), (3, 1)]:
            got = emu_refine(emu, cm, nlevels, blocks, reverse)
            assert same_mesh(got, want), (name, nlevels, blocks, reverse)
        assert np.isfinite(want[0]).all()


def test_random_campaign(emu):
    rs = np.random.RandomState(2024)
    cases = 0
    for it in range(120):
        kind = rs.randint(4)
        if kind == 0:
            mesh = grid(int(rs.randint(1, 9)), int(rs.randint(1, 9)), rs, holes=float(rs.choice([0, 0, .1, .3])))
        elif kind == 1:
            mesh = fan(int(rs.randint(3, 20)), True, rs, centre_z=float(rs.rand()))
        elif kind == 2:
            mesh = fan(int(rs.randint(1, 15)), False, rs)
        else:
            mesh = closed_solid("oct", rs)
        if not mesh[0]:
            continue
        cm = pb.ControlMesh(*mesh)
        for nlevels in (1, 2, 3):
            got = emu_refine(emu, cm, nlevels, int(rs.choice([0, 1, 2])), int(rs.randint(2)))
            assert same_mesh(got, cm.host_refine(nlevels)), (it, kind, nlevels)
            cases += 1
    assert cases > 300


def _killeroo_control_mesh():
    from pbrt_v2_b200 import scenes
    scenes.stage()
    path = os.path.join(scenes.STAGED, "geometry", "killeroo.pbrt")
    if not os.path.exists(path):
        pytest.skip("fixture scenes are not staged")
    text = open(path).read()
    m = re.search(r'Shape "loopsubdiv".*?"integer indices"\s*\[([^\]]*)\]\s*"point P"\s*\[([^\]]*)\]', text, re.S)
    assert m
    return np.array(m.group(1).split(), np.int32).reshape(-1, 3), np.array(m.group(2).split(), np.float32).reshape(-1, 3)


def test_killeroo_control_mesh(emu):
    """the loopsubdiv shape of scenes/geometry/killeroo.pbrt (BASELINE configs C1, C3, C4) at the scene's nlevels and one more"""
    cm = pb.ControlMesh(*_killeroo_control_mesh())
    assert len(cm.indices) > 1000
    for nlevels in (1, 2):
        assert same_mesh(emu_refine(emu, cm, nlevels), cm.host_refine(nlevels))


def test_entry_point_validates_before_touching_the_device():
    tris, P = grid(3, 3, np.random.RandomState(1))
    good = pb.ControlMesh(tris, P)
    assert good.counts(2)[1] == len(tris) * 16
    # an edge in three faces (the reference's own edge map mis-links such a mesh and its constructor may not terminate,
    # so the descriptor is put together by hand): refused
    three = pb.ControlMesh([[0, 1, 2], [1, 0, 3], [5, 6, 4]], np.random.RandomState(2).rand(7, 3))
    three.faceV[2] = [0, 1, 4]
    with pytest.raises(pb.PbrtGpuError, match="more than two faces"):
        three.counts(1)
    bad = pb.ControlMesh(tris, P)
    bad.faceF[0, 0] = 5 if bad.faceF[0, 0] != 5 else 6
    with pytest.raises(pb.PbrtGpuError, match="reciprocal|neighbour"):
        bad.counts(1)
    bad = pb.ControlMesh(tris, P)
    bad.faceV[1, 1] = 10 ** 6
    with pytest.raises(pb.PbrtGpuError, match="bad vertex index"):
        bad.counts(1)
    nanP = P.copy()
    nanP[2, 1] = np.nan
    with pytest.raises(pb.PbrtGpuError, match="not finite"):
        pb.ControlMesh(tris, nanP).counts(1)
    with pytest.raises(pb.PbrtGpuError, match="nLevels"):
        good.counts(13)
    if pb.device_count() == 0:
        with pytest.raises(pb.PbrtGpuError, match="no CPU fallback"):
            good.refine(1)


def test_front_end_refiner_option_has_no_silent_fallback(staged, monkeypatch):
    """Shape "loopsubdiv" "string refiner" "gpu" (or PBRTGPU_SUBDIV=gpu) refines on the device or fails the scene load"""
    from helpers import host_scene
    if pb.device_count() > 0:
        pytest.skip("a CUDA device is present: covered by test_zz_gpu_loopsubdiv.py")
    monkeypatch.setenv("PBRTGPU_SUBDIV", "gpu")
    with pytest.raises(pb.PbrtGpuError, match="no CUDA device"):
        host_scene("killeroo-simple", 24, 24, 4, 2, tag="-devsubdiv-cpu")
