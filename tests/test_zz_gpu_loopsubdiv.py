"""pbrtgpu_loopsubdiv_refine on the GPU against the host front end's LoopSubdiv::Refine: the same
mesh, vertex for vertex; and a scene whose loopsubdiv shapes were refined on the device renders the same film.
(Named to sort last: the kernels were first verified on the CPU -- test_loopsubdiv_on_host.py -- and these are their
first runs on hardware.)"""
import numpy as np
import pytest

import pbrt_v2_b200 as pb
from helpers import host_scene
from test_loopsubdiv_on_host import _killeroo_control_mesh, control_meshes, same_mesh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,mesh", control_meshes(), ids=[n for n, _ in control_meshes()])
def test_device_refine_matches_host_refine(name, mesh):
    cm = pb.ControlMesh(*mesh)
    for nlevels in (0, 1, 2, 3, 5):
        assert same_mesh(cm.refine(nlevels), cm.host_refine(nlevels)), (name, nlevels)


def test_killeroo_control_mesh_and_timing():
    import time
    cm = pb.ControlMesh(*_killeroo_control_mesh())
    cm.refine(0)   # context creation outside the timing
    for nlevels in (1, 3, 4):
        t0 = time.time()
        want = cm.host_refine(nlevels)
        t1 = time.time()
        got = cm.refine(nlevels)
        t2 = time.time()
        assert same_mesh(got, want), nlevels
        print("\nLoopSubdiv killeroo nlevels %d: %d triangles, host %.3f s, device call %.3f s" % (nlevels, len(want[2]), t1 - t0, t2 - t1))


def test_scene_refined_on_device_is_the_same_scene(staged, monkeypatch):
    host = host_scene("killeroo-simple", 48, 48, 4, 8)
    monkeypatch.setenv("PBRTGPU_SUBDIV", "gpu")
    dev = host_scene("killeroo-simple", 48, 48, 4, 8, tag="-devsubdiv")
    monkeypatch.delenv("PBRTGPU_SUBDIV")
    assert dev is not host and dev.counts() == host.counts()
    assert dev.nodes().tobytes() == host.nodes().tobytes()
    px_h, _ = host.render()
    px_d, _ = dev.render()
    assert px_h.tobytes() == px_d.tobytes()
