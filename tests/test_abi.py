"""C-ABI checks that need no GPU: the shared libraries load, export every entry
point include/pbrtgpu.h declares, agree with the Python mirrors on struct layout,
and refuse to compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

import pbrt_v2_b200 as pb

HEADER = os.path.join(pb.REPO, "include", "pbrtgpu.h")


def _declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pbrtgpu_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = pb.gpu_lib()
    names = _declared_functions()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), "libpbrtgpu.so does not export %s" % n


def test_struct_layouts_match_header():
    src = r'''
    #include <stdio.h>
    #include <stddef.h>
    #include "pbrtgpu.h"
    int main(void) {
      printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(PbrtGpuRenderDesc), sizeof(PbrtGpuStats),
             sizeof(PbrtGpuPixel), sizeof(PbrtGpuRay), sizeof(PbrtGpuHit), sizeof(PbrtGpuBVHNode),
             sizeof(PbrtGpuSceneDesc), offsetof(PbrtGpuRenderDesc, nTasks), offsetof(PbrtGpuRenderDesc, shardCount),
             sizeof(PbrtGpuMaterial));
      return 0; }'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "t")
        subprocess.run(["gcc", "-I", os.path.dirname(HEADER), c, "-o", exe], check=True)   # the header is plain C
        v = [int(x) for x in subprocess.run([exe], check=True, stdout=subprocess.PIPE, text=True).stdout.split()]
    assert v[0] == C.sizeof(pb.RenderDesc)
    assert v[1] == C.sizeof(pb.Stats)
    assert v[2] == pb.PIXEL_DTYPE.itemsize == 16
    assert v[3] == pb.RAY_DTYPE.itemsize
    assert v[4] == pb.HIT_DTYPE.itemsize
    assert v[5] == 32, "PbrtGpuBVHNode must keep the reference's 32-byte LinearBVHNode layout"
    assert v[7] == pb.RenderDesc.nTasks.offset
    assert v[8] == pb.RenderDesc.shardCount.offset


def test_no_cpu_fallback(staged):
    """Without a CUDA device every compute entry point must fail loudly (PBRTGPU_ENODEVICE)."""
    if pb.device_count() > 0:
        pytest.skip("a GPU is present")
    from helpers import host_scene
    hs = host_scene("bunny", 16, 16, 4, 8)
    with pytest.raises(pb.PbrtGpuError):
        pb.GpuScene(hs)
    with pytest.raises(pb.PbrtGpuError):
        pb.ld_samples(hs.render_desc(), 0)
    with pytest.raises(pb.PbrtGpuError):
        hs.render()


def test_product_never_loads_the_oracle():
    """The product package must not reference oracle/ in any way."""
    pkg = os.path.join(pb.REPO, "pbrt_v2_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".h", ".cu", ".cuh")):
                text = open(os.path.join(root, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle_api" not in text and "oracle/" not in text.replace("(oracle/", "").replace("``oracle/``", ""), \
                    "%s mentions the oracle" % os.path.join(root, f)


def _rd(res=64, spp=4, ntasks=1024):
    rd = pb.set_path_layout(pb.RenderDesc())
    rd.xResolution = rd.yResolution = res
    rd.cropWindow[:] = [0, 1, 0, 1]
    rd.filterXWidth = rd.filterYWidth = 0.5
    rd.samplesPerPixel = spp
    rd.maxDepth = 5
    rd.nTasks, rd.taskBegin, rd.taskEnd = ntasks, 0, ntasks
    rd.shutterOpen, rd.shutterClose = 0.0, 1.0
    return rd


def test_render_desc_validation_precedes_the_device():
    """Descriptor errors are reported as such (EINVAL / EUNSUPPORTED) even on a box without a GPU."""
    cases = []
    rd = _rd(res=4096, ntasks=1 << 24)            # dy * nx overflows int in Sampler::ComputeSubWindow (sampler.cpp:58)
    cases.append((rd, "ComputeSubWindow"))
    rd = _rd(); rd.size1D[3] = 2                  # not PathIntegrator's Sample layout
    cases.append((rd, "Sample layout"))
    rd = _rd(spp=3)                               # LDSampler rounds to powers of two
    cases.append((rd, "power of two"))
    rd = _rd(); rd.integrator = 7
    cases.append((rd, "integrator"))
    rd = _rd(); rd.filterXWidth = 2.0             # wide filters need cross-task splats
    cases.append((rd, "filters wider"))
    rd = _rd(); rd.cropWindow[:] = [0.5, 0.25, 0, 1]   # low > high
    cases.append((rd, "crop window"))
    rd = _rd(); rd.cropWindow[3] = float("nan")
    cases.append((rd, "crop window"))
    rd = _rd(); rd.shutterClose = float("inf")   # every ray's time comes from the shutter interval
    cases.append((rd, "non-finite camera"))
    rd = _rd(res=65536)
    cases.append((rd, "2^28 pixels"))
    for rd, needle in cases:
        with pytest.raises(pb.PbrtGpuError) as e:
            pb.ld_samples(rd, 0)
        assert needle in str(e.value), str(e.value)


def test_render_kernels_are_the_gpu_verified_ones():
    """No GPU here: what guards the kernels between GPU visits is their SASS.  scripts/sass_fingerprint.sh holds the md5 of
    pbrtgpu.cu's cubin as it ran the round's last full GPU visit (profiles/r02g_*); experiments that are meant to be
    compiled out (-DPBRT_METAL=0 ...) must not move it.  (A deliberate kernel change updates the fingerprint together with a
    GPU run.)"""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not installed")
    r = subprocess.run(["bash", os.path.join(pb.REPO, "scripts", "sass_fingerprint.sh")], cwd=pb.REPO, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def test_scene_validate_refuses_malformed_descriptors(staged):
    """pbrtgpu_scene_validate (the checks of pbrtgpu_scene_create, no device needed): a front-end descriptor passes; indices
    past their arrays, a BVH whose second child does not follow the first child's subtree, a BVH deeper than the 64-entry
    traversal stack are PBRTGPU_EINVAL / PBRTGPU_EUNSUPPORTED (scripts/fuzz_descriptors.py is the broad version of this)"""
    import ctypes as C
    from helpers import host_scene
    hs = host_scene("bunny", 16, 16, 1, 1)
    assert pb.scene_validate(hs.desc_ptr)[0] == 0
    src = pb.SceneDesc.from_address(hs.desc_ptr)

    def mutated(field, count, dtype, edit):
        d = pb.SceneDesc()
        C.memmove(C.byref(d), hs.desc_ptr, C.sizeof(pb.SceneDesc))
        a = np.empty(getattr(src, count), dtype)
        C.memmove(a.ctypes.data, getattr(src, field), a.nbytes)
        edit(a)
        setattr(d, field, a.ctypes.data)
        rc, msg = pb.scene_validate(C.addressof(d))
        return rc, msg

    tri = np.dtype([("v", np.uint32, 3), ("mesh", np.uint32)])
    rc, msg = mutated("triangles", "nTriangles", tri, lambda a: a["v"].__setitem__((5, 1), src.nVertices))
    assert rc != 0 and "vertex index" in msg
    rc, msg = mutated("triangles", "nTriangles", tri, lambda a: a["mesh"].__setitem__(7, 99))
    assert rc != 0 and "mesh index" in msg

    def second_child_first(a):
        i = int(np.nonzero(a["nPrimitives"] == 0)[0][3])
        a["offset"][i] = i + 1
    rc, msg = mutated("nodes", "nNodes", pb.NODE_DTYPE, second_child_first)
    assert rc != 0 and "second child" in msg

    def chain(a):   # a 70-level comb: every interior node's first child is the next interior node
        n = len(a)
        for i in range(70):
            a["nPrimitives"][i] = 0
            a["axis"][i] = 0
            a["offset"][i] = n - 1 - i      # second child: some later node
        for i in range(n - 71, n):
            a["nPrimitives"][i] = 1
            a["offset"][i] = 0
    rc, msg = mutated("nodes", "nNodes", pb.NODE_DTYPE, chain)
    assert rc != 0 and "deeper than 64" in msg
