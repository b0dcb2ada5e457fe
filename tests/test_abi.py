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
