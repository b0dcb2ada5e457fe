"""GPU parity of the specular-material device code, which ships DISABLED (compile flag PBRT_SPECULAR, see
device/dshade.cuh): these tests run only when PBRTGPU_LIB points at a library built with it, e.g.

    make NVDEFS=-DPBRT_SPECULAR=1 ...   or   scripts/r2_specular.sh   (builds pbrt_v2_b200/ab/libpbrtgpu_specular.so)
    PBRTGPU_LIB=ab/libpbrtgpu_specular.so python -m pytest tests/test_gpu_specular.py -m gpu

With the shipped library they are skipped (and test_gpu_parity.py checks that it refuses the materials)."""
import os

import numpy as np
import pytest

import pbrt_v2_b200 as pb
import oracle_api as oa
from helpers import variant_scene

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif("specular" not in os.environ.get("PBRTGPU_LIB", ""), reason="needs a PBRT_SPECULAR build (PBRTGPU_LIB)")]


@pytest.mark.parametrize("name", ["path-glass-mirror", "path-glass-params-mirror-floor"])
def test_specular_path_integrator_bit_exact_vs_oracle(staged, name):
    hs, _ = variant_scene(name)
    px_gpu, _ = pb.GpuScene(hs).render()
    px_cpu, _ = oa.render(hs, threads=1)
    assert np.array_equal(px_gpu["weightSum"], px_cpu["weightSum"])
    own = px_cpu["weightSum"] == hs.render_desc().samplesPerPixel
    assert np.array_equal(px_gpu["Lxyz"][own].view(np.uint32), px_cpu["Lxyz"][own].view(np.uint32))


def test_specular_under_directlighting_is_still_refused(staged):
    hs, _ = variant_scene("direct-glass-mirror")
    gs = pb.GpuScene(hs)
    with pytest.raises(pb.PbrtGpuError) as e:
        gs.render()
    assert "SpecularReflect" in str(e.value)
