"""No silent fallbacks: what the oracle and the host front end can do but the kernels cannot yet (SURVEY 8(f) rank 4 rest)
must be refused by the device entry points with PBRTGPU_EUNSUPPORTED, not rendered differently."""
import pytest

import pbrt_v2_b200 as pb
from helpers import variant_scene

pytestmark = pytest.mark.gpu


def test_infinite_area_light_is_refused(staged):
    hs, _ = variant_scene("path-infinite")
    with pytest.raises(pb.PbrtGpuError) as e:
        pb.GpuScene(hs)
    assert "infinite area lights" in str(e.value)


@pytest.mark.parametrize("name", ["filter-gaussian", "filter-widebox"])
def test_wide_pixel_filters_are_refused(staged, name):
    hs, _ = variant_scene(name)
    gs = pb.GpuScene(hs)           # the scene itself is fine
    with pytest.raises(pb.PbrtGpuError) as e:
        gs.render()
    assert "filters" in str(e.value)
