"""GPU parity at the REAL settings of BASELINE.json configs[1..4].

Render tasks are independent (own MT19937 stream seeded with the task number, own pixels:
samplerrenderer.cpp:60-164, :203-212), so a sub-range [taskBegin, taskEnd) of the full-size job is
exactly what the full render computes for those pixels.  Each case renders a few sub-ranges of >= 256
tasks -- first tasks, last tasks, an odd window over the object -- at the configuration's full
resolution, spp and task grid, on the GPU through the C ABI and on the CPU oracle, and compares them
BIT FOR BIT on every pixel whose samples all come from its own task (weightSum == spp; a pixel that
received an edge splat from a neighbouring task is order-dependent in the reference itself).

What this reaches that the small cases do not: the 4096-spp grouped-shuffle table path (C3,
lowdiscrepancy.cpp:67-79, montecarlo.cpp:200-252), `nlevels 3` nested BVHs under TransformedPrimitives
(C4), multi-pixel tasks that run in 64 rounds over >= 1 M triangles at maxdepth 8 (C5).
"""
import numpy as np
import pytest

import pbrt_v2_b200 as pb
import oracle_api as oa
from helpers import host_scene

pytestmark = pytest.mark.gpu


def _ncores(xres, yres):
    import bench
    return bench.per_pixel_ncores(xres, yres)


def _task_of_pixel(rd, x, y):
    """the task whose window contains pixel (x, y): linear search over the oracle's ComputeSubWindow restatement
    around the proportional guess (tile borders are floor(lerp) values, sampler.cpp:55-74)"""
    n = rd.nTasks
    nx, ny = n, 1
    while nx % 2 == 0 and 2 * rd.xResolution * ny < rd.yResolution * nx:
        nx //= 2
        ny *= 2
    gx, gy = min(nx - 1, x * nx // rd.xResolution), min(ny - 1, y * ny // rd.yResolution)
    for dy in (0, -1, 1):
        for dx in (0, -1, 1):
            t = (gy + dy) * nx + gx + dx
            if 0 <= t < n:
                x0, x1, y0, y1 = oa.task_window(rd, t)
                if x0 <= x < x1 and y0 <= y < y1:
                    return t
    raise AssertionError("no task owns pixel (%d, %d)" % (x, y))


def _compare_subrange(hs, gs, rd, begin, end, spp, min_own=0.25):
    rd.taskBegin, rd.taskEnd = int(begin), int(end)
    px_gpu, st = gs.render(rd)
    px_cpu, cnt = oa.render(hs, rd)
    assert st.cameraSamples == cnt.cameraSamples and cnt.cameraSamples > 0
    if not np.array_equal(px_gpu["weightSum"], px_cpu["weightSum"]):
        ys, xs = np.nonzero(px_gpu["weightSum"] != px_cpu["weightSum"])
        raise AssertionError("sample counts differ on %d pixels in tasks [%d, %d): first (x, y, gpu, cpu) = %s" % (
            len(xs), begin, end, [(int(x), int(y), float(px_gpu["weightSum"][y, x]), float(px_cpu["weightSum"][y, x]))
                                  for x, y in list(zip(xs, ys))[:6]]))
    touched = px_cpu["weightSum"] > 0
    own = px_cpu["weightSum"] == spp
    assert own.sum() >= min_own * touched.sum()
    a, b = px_gpu["Lxyz"][own].view(np.uint32), px_cpu["Lxyz"][own].view(np.uint32)
    bad = (a != b).any(axis=1)
    assert not bad.any(), "%d of %d fully-owned pixels differ bitwise in tasks [%d, %d)" % (bad.sum(), own.sum(), begin, end)
    # the rest (edge-splat receivers) still has to agree to float tolerance
    rest = touched & ~own
    if rest.any():
        np.testing.assert_allclose(px_gpu["Lxyz"][rest], px_cpu["Lxyz"][rest], rtol=2e-4, atol=1e-6)
    return int(own.sum()), float(px_cpu["Lxyz"][own].sum())


def _ranges(rd, hot_pixel, n=256):
    """first tasks, last tasks, and an odd-aligned window around the task that owns `hot_pixel` (object centre)"""
    t = _task_of_pixel(rd, *hot_pixel)
    mid = max(0, min(rd.nTasks - n - 3, t - n // 2 + 1)) | 1   # odd start
    return [(0, n), (mid, mid + n + 3), (rd.nTasks - n, rd.nTasks)]


def test_c2_bunny_1024sq_1024spp(staged):
    """configs[1]: scenes/bunny.pbrt, path, 1024 spp, 1024x1024: one task per pixel (2^20 tasks)"""
    hs = host_scene("bunny", 1024, 1024, 1024, _ncores(1024, 1024))
    rd = hs.render_desc()
    assert rd.nTasks == 1 << 20 and rd.samplesPerPixel == 1024
    gs = pb.GpuScene(hs)
    lit = 0.0
    for b, e in _ranges(rd, (512, 600)):
        n, s = _compare_subrange(hs, gs, rd, b, e, 1024)
        lit += s
    assert lit > 0, "the sub-ranges never saw the bunny"


def test_c3_killeroo_1080p_4096spp(staged):
    """configs[2]: scenes/killeroo-simple.pbrt, path, 4096 spp (grouped table shuffles), 1920x1080, 2^20-task grid"""
    hs = host_scene("killeroo-simple", 1920, 1080, 4096, _ncores(1920, 1080))
    rd = hs.render_desc()
    assert rd.samplesPerPixel == 4096 and rd.nTasks == 1 << 20
    gs = pb.GpuScene(hs)
    lit = 0.0
    for b, e in _ranges(rd, (960, 600)):
        n, s = _compare_subrange(hs, gs, rd, b, e, 4096)
        lit += s
    assert lit > 0


def test_c4_anim_nlevels3_1080p_1024spp(staged):
    """configs[3]: scenes/anim-killeroos-moving.pbrt with geometry/killeroo.pbrt at nlevels 3 (2 x 532 224 triangles in
    nested BVHs under animated TransformedPrimitives), path, 1024 spp, 1920x1080"""
    hs = host_scene("anim-killeroos-moving", 1920, 1080, 1024, _ncores(1920, 1080), nlevels=3)
    assert hs.counts()["triangles"] >= 1000000
    rd = hs.render_desc()
    gs = pb.GpuScene(hs)
    lit = 0.0
    for b, e in _ranges(rd, (960, 560)):
        n, s = _compare_subrange(hs, gs, rd, b, e, 1024)
        lit += s
    assert lit > 0


def test_c5_synthetic_4096sq_512spp(staged):
    """configs[4] at >= 1 M triangles: synthetic random mesh, path maxdepth 8, 512 spp, 4096x4096: 2^18 tasks of
    8x8 pixels (64 rounds per task: the RNG stream runs on from pixel to pixel)"""
    hs = pb.HostScene.synthetic(1000000, 4096, 4096, 512, maxdepth=8, ncores=_ncores(4096, 4096))
    rd = hs.render_desc()
    assert rd.nTasks == 1 << 18 and rd.maxDepth == 8
    x0, x1, y0, y1 = oa.task_window(rd, rd.nTasks // 2 + 7)
    assert (x1 - x0) * (y1 - y0) == 64
    gs = pb.GpuScene(hs)
    lit = 0.0
    for b, e in _ranges(rd, (2048, 2048)):
        n, s = _compare_subrange(hs, gs, rd, b, e, 512)
        assert n >= 256 * 64 * 0.25
        lit += s
    assert lit > 0
