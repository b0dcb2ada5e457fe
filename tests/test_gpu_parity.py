"""GPU parity tests (run on the B200 box): the CUDA hot path, called through the
C ABI of include/pbrtgpu.h, against the CPU oracle on the same inputs.

Bit-exact where the arithmetic is integer or IEEE float32/float64 (+,-,*,/,sqrt:
RNG streams, LD sample tables, BVH traversal, triangle hits); for rendered pixels
the tolerance below applies because the device evaluates sinf/cosf/powf/expf/
atan2f/acosf as correctly-rounded double results, which differ from glibc's by
<= 1 ulp on a small share of arguments (DESIGN.md "libm").
"""
import numpy as np
import pytest

import pbrt_v2_b200 as pb
import oracle_api as oa
from helpers import host_scene, random_rays, scene_bounds

pytestmark = pytest.mark.gpu

# per-pixel relative tolerance on RGB for stream-exact pixels (SURVEY 7.3 Tier A)
TIER_A_REL = 2e-4


def _gpu_scene(hs):
    return pb.GpuScene(hs)


@pytest.mark.parametrize("spp,ncores", [(4, 32), (64, 32), (1024, 2)])
def test_ld_samples_bit_exact(staged, spp, ncores):
    hs = host_scene("bunny", 32, 32, spp, ncores)
    rd = hs.render_desc()
    for task in (0, 1, rd.nTasks // 2 + 3, rd.nTasks - 1):
        if oa.task_window(rd, task)[0] == oa.task_window(rd, task)[1]:
            continue
        want = oa.ld_samples(rd, task)
        got = pb.ld_samples(rd, task)
        assert np.array_equal(want.view(np.uint32), got.view(np.uint32)), "LD tables differ for task %d" % task


def test_ld_samples_general_layouts_bit_exact():
    """k_tables_general (arrays with several values per pixel sample: DirectLightingIntegrator's layouts) against
    the vectors minted from the reference's own Sample / LDSampler classes, and against the oracle."""
    import os
    from helpers import layout_rd
    g = np.load(os.path.join(pb.REPO, "tests", "golden", "ld_samples_general.npz"))
    for key in [k for k in g.files if not k.endswith("_next") and not k.endswith("_meta")]:
        rd, task = layout_rd(g[key + "_meta"])
        got = pb.ld_samples(rd, task)
        assert got.shape == g[key].shape
        assert np.array_equal(got.view(np.uint32), g[key].view(np.uint32)), key
        for t in (0, rd.nTasks - 1):
            assert np.array_equal(pb.ld_samples(rd, t).view(np.uint32), oa.ld_samples(rd, t).view(np.uint32))


@pytest.mark.parametrize("base", ["bunny", "killeroo-simple", "anim-killeroos-moving"])
def test_intersect_bit_exact(staged, base):
    hs = host_scene(base, 32, 32, 4, 32)
    lo, hi = scene_bounds(hs)
    if base.startswith("anim"):   # the scene's floor planes are 2000 units wide: aim at the moving killeroos
        lo, hi = np.maximum(lo, [-300, -300, -150]).astype(np.float32), np.minimum(hi, [300, 300, 150]).astype(np.float32)
    rays = random_rays(200000, 7, lo, hi)
    # ray.time: inside, at and beyond both ends of the [0,1] transform interval (AnimatedTransform::Interpolate's cases)
    rays["time"] = np.random.RandomState(11).choice(np.array([-0.5, 0.0, 0.25, 0.5, 0.999, 1.0, 1.5], np.float32), len(rays))
    rays["time"][::3] = np.random.RandomState(12).rand(len(rays[::3])).astype(np.float32)
    gs = _gpu_scene(hs)
    got = gs.intersect(rays)
    want = oa.intersect(hs, rays)
    assert np.array_equal(got["hit"], want["hit"])
    assert np.array_equal(got["t"].view(np.uint32), want["t"].view(np.uint32))
    h = want["hit"] == 1
    assert np.array_equal(got["primitive"][h], want["primitive"][h])
    assert np.array_equal(got["b1"][h].view(np.uint32), want["b1"][h].view(np.uint32))
    # dg.p / dg.nn: exact for triangles; quadric hits go through atan2f/acosf/sinf
    tri = h & (np.abs(want["b1"]) + np.abs(want["b2"]) > 0)
    assert np.array_equal(got["p"][tri].view(np.uint32), want["p"][tri].view(np.uint32))
    assert np.array_equal(got["n"][tri].view(np.uint32), want["n"][tri].view(np.uint32))
    np.testing.assert_allclose(got["n"][h], want["n"][h], atol=2e-6)
    occ_got = gs.intersect_p(rays)
    occ_want = oa.intersect_p(hs, rays)
    assert np.array_equal(occ_got, occ_want)
    assert h.mean() > 0.2, "ray set should hit geometry reasonably often"
    if base.startswith("anim"):
        nested = h & (want["primitive"] >= 7)   # the scene BVH holds 7 primitives; the rest live under a TransformedPrimitive
        assert nested.mean() > 0.02, "ray set should reach the animated killeroos"
        assert np.array_equal(got["p"][nested].view(np.uint32), want["p"][nested].view(np.uint32))
        assert np.array_equal(got["n"][nested].view(np.uint32), want["n"][nested].view(np.uint32))


def _compare_images(hs, px_gpu, px_cpu):
    rgb_g, rgb_c = hs.pixels_to_rgb(px_gpu), hs.pixels_to_rgb(px_cpu)
    assert np.array_equal(px_gpu["weightSum"], px_cpu["weightSum"]), "sample counts per pixel differ"
    denom = np.maximum(np.abs(rgb_c), 1e-3)
    rel = np.abs(rgb_g - rgb_c) / denom
    bad = (rel > TIER_A_REL).any(axis=2)
    return rel, bad


@pytest.mark.parametrize("base,res,spp,ncores", [
    ("bunny", 64, 16, 128),            # one task per pixel (nTasks = 4096)
    ("bunny", 64, 16, 1),              # 32 tasks of 128 pixels: the sequential-stream case
    ("killeroo-simple", 64, 16, 128),
    ("killeroo-simple", 48, 8, 2),
])
def test_render_matches_oracle(staged, base, res, spp, ncores):
    hs = host_scene(base, res, res, spp, ncores)
    gs = _gpu_scene(hs)
    px_gpu, st = gs.render()
    px_cpu, cnt = oa.render(hs)
    rel, bad = _compare_images(hs, px_gpu, px_cpu)
    # Tier A: stream-exact pixels agree to TIER_A_REL; Tier B: a libm-ulp flip of one
    # hit/miss or Russian-roulette decision re-randomises the rest of that task
    assert bad.mean() <= 0.01, "%.2f%% of pixels differ by more than %g" % (100 * bad.mean(), TIER_A_REL)
    rgb_g, rgb_c = hs.pixels_to_rgb(px_gpu), hs.pixels_to_rgb(px_cpu)
    assert abs(rgb_g.mean() - rgb_c.mean()) <= 1e-3 * abs(rgb_c.mean())
    assert st.cameraSamples == cnt.cameraSamples


@pytest.mark.parametrize("base,res,spp,ncores,integrator,strategy", [
    ("bunny", 64, 16, 128, "path", None), ("bunny", 64, 16, 1, "path", None),
    ("killeroo-simple", 64, 16, 128, "path", None), ("killeroo-simple", 48, 8, 2, "path", None),
    ("anim-killeroos-moving", 64, 16, 128, "path", None), ("anim-killeroos-moving", 48, 8, 2, "path", None),   # config C4
    # DirectLightingIntegrator: strategies "all" (per-light sample arrays) and "one", per-pixel and multi-pixel tasks
    ("killeroo-simple", 64, 16, 128, "directlighting", None), ("killeroo-simple", 48, 8, 2, "directlighting", None),
    ("bunny", 64, 8, 128, "directlighting", None), ("bunny", 48, 16, 1, "directlighting", None),
    ("bunny", 48, 8, 2, "directlighting", "one"), ("anim-killeroos-moving", 48, 8, 2, "directlighting", None),
])
def test_render_bit_exact_vs_oracle(staged, base, res, spp, ncores, integrator, strategy):
    """With glibc's libm restated on the device (refmath.h) the film is bit-identical to the oracle's, except
    (a) pixels that receive an edge splat from another task (order-dependent in the reference itself) and
    (b) the ~1e-9 of expf arguments where the host libm's FMA build rounds differently."""
    hs = host_scene(base, res, res, spp, ncores, integrator=integrator, strategy=strategy)
    px_gpu, _ = _gpu_scene(hs).render()
    px_cpu, _ = oa.render(hs, threads=1)
    assert np.array_equal(px_gpu["weightSum"], px_cpu["weightSum"])
    diff = (px_gpu["Lxyz"].view(np.uint32) != px_cpu["Lxyz"].view(np.uint32)).any(axis=2)
    splat = px_cpu["weightSum"] != spp
    unexplained = diff & ~splat
    print("differing pixels: %d (of which splat receivers %d) / %d" % (diff.sum(), (diff & splat).sum(), diff.size))
    assert unexplained.sum() <= max(1, diff.size // 2000), "%d pixels differ bitwise without being splat receivers" % unexplained.sum()
    np.testing.assert_allclose(px_gpu["Lxyz"][~splat & ~diff], px_cpu["Lxyz"][~splat & ~diff], rtol=0, atol=0)


@pytest.mark.parametrize("name", ["dof", "maxdepth1", "maxdepth8", "nonsquare-crop", "directlighting-maxdepth1", "shutter",
                                  "spp1", "film1x1", "maxdepth0", "nolights", "nolights-directlighting"])
def test_render_variants_bit_exact_vs_oracle(staged, name):
    """the scene-file variants the oracle is pinned on against the reference binary (tests/helpers.py VARIANTS)"""
    from helpers import variant_scene
    hs, _ = variant_scene(name)
    px_gpu, _ = _gpu_scene(hs).render()
    px_cpu, _ = oa.render(hs, threads=1)
    assert np.array_equal(px_gpu["weightSum"], px_cpu["weightSum"])
    own = px_cpu["weightSum"] == hs.render_desc().samplesPerPixel   # (pixels that received an edge splat: order-dependent)
    assert own.mean() > 0.99 or own.size == 1
    assert np.array_equal(px_gpu["Lxyz"][own].view(np.uint32), px_cpu["Lxyz"][own].view(np.uint32))


def test_task_sharding_is_exact(staged):
    """rendering [0,n/2) and [n/2,n) separately sums to the one-shot render (multi-GPU split)."""
    hs = host_scene("bunny", 48, 48, 8, 72)   # nTasks = 4096 > 2304 pixels: some tasks are empty
    gs = _gpu_scene(hs)
    rd = hs.render_desc()
    full, _ = gs.render(rd)
    half = rd.nTasks // 2
    rd.taskBegin, rd.taskEnd = 0, half
    a, _ = gs.render(rd)
    rd.taskBegin, rd.taskEnd = half, rd.nTasks
    b, _ = gs.render(rd)
    s = a["weightSum"] + b["weightSum"]
    assert np.array_equal(s, full["weightSum"])
    # pixels owned by exactly one shard are bit-identical to the full render
    only_a = (b["weightSum"] == 0)
    assert np.array_equal(a["Lxyz"][only_a].view(np.uint32), full["Lxyz"][only_a].view(np.uint32))
    only_b = (a["weightSum"] == 0)
    assert np.array_equal(b["Lxyz"][only_b].view(np.uint32), full["Lxyz"][only_b].view(np.uint32))


def test_lanes_do_not_change_the_film(staged, monkeypatch):
    """The render call deals its task batches to concurrent lanes (streams + host threads); batches share
    nothing, so the film must not depend on the number of lanes (splat-receiving pixels aside)."""
    hs = host_scene("killeroo-simple", 96, 96, 8, 288)   # 16384 tasks: enough for 4 lanes
    gs = _gpu_scene(hs)
    films = []
    for lanes in ("1", "2", "4"):
        monkeypatch.setenv("PBRTGPU_LANES", lanes)
        monkeypatch.setenv("PBRTGPU_SCRATCH_GB", "0.25")   # several batches per lane
        px, st = gs.render()
        films.append(px)
    own = films[0]["weightSum"] == 8
    for f in films[1:]:
        assert np.array_equal(f["weightSum"], films[0]["weightSum"])
        assert np.array_equal(f["Lxyz"][own].view(np.uint32), films[0]["Lxyz"][own].view(np.uint32))


def test_scene_update_reuses_the_handle(staged):
    """pbrtgpu_scene_update: one handle, two different scenes in turn; films equal fresh handles' films."""
    a, b = host_scene("bunny", 32, 32, 4, 32), host_scene("killeroo-simple", 32, 32, 4, 32)
    want_a, _ = _gpu_scene(a).render()
    want_b, _ = _gpu_scene(b).render()
    gs = pb.GpuScene(a)
    for hs, want in ((b, want_b), (a, want_a), (b, want_b)):
        gs.update(hs)
        got, _ = gs.render(hs.render_desc())
        assert np.array_equal(got["Lxyz"].view(np.uint32), want["Lxyz"].view(np.uint32))
    gs.close()


def test_host_renderer_entry_point(staged):
    """GpuSamplerRenderer::Render (C++ host class -> C ABI) gives the same film as the direct call."""
    hs = host_scene("bunny", 32, 32, 4, 32)
    px1, st = hs.render()
    gs = _gpu_scene(hs)
    px2, _ = gs.render()
    assert np.array_equal(px1["Lxyz"].view(np.uint32), px2["Lxyz"].view(np.uint32))
    assert st.cameraSamples == 32 * 32 * 4


@pytest.mark.parametrize("key", ["bunny_r32_s4_c32", "bunny_r32_s8_c1", "killeroo_simple_r32_s4_c32",
                                 "killeroo_simple_r32_s8_c1", "anim_killeroos_moving_r32_s4_c32",
                                 "anim_killeroos_moving_r32_s8_c1",
                                 "directlighting_x_killeroo_simple_r32_s4_c32", "directlighting_x_killeroo_simple_r32_s8_c1",
                                 "directlighting_x_bunny_r32_s4_c32", "directlightingone_x_bunny_r32_s8_c1",
                                 "directlighting_x_anim_killeroos_moving_r32_s4_c1"])
def test_render_matches_golden_reference_images(staged, key):
    """GPU render vs images rendered by the UNMODIFIED reference (tests/golden/images.npz)."""
    import os
    g = np.load(os.path.join(pb.REPO, "tests", "golden", "images.npz"))
    from helpers import golden_image_case
    base, res, spp, ncores, integrator, strategy = golden_image_case(key)
    hs = host_scene(base, res, res, spp, ncores, integrator=integrator, strategy=strategy)
    px, _ = _gpu_scene(hs).render()
    rgb, want = hs.pixels_to_rgb(px), g[key]
    rel = np.abs(rgb - want) / np.maximum(np.abs(want), 1e-3)
    bad = (rel > TIER_A_REL).any(axis=2)
    assert bad.mean() <= 0.01, "%.2f%% of pixels differ from the reference by more than %g" % (100 * bad.mean(), TIER_A_REL)
    assert abs(rgb.mean() - want.mean()) <= 1e-3 * abs(want.mean())


def test_full_size_properties(staged):
    """BASELINE-size film (1024x1024) at reduced spp: size-independent properties --
    every pixel receives exactly spp samples (+ edge splats), shards tile the image, no NaN/negative."""
    hs = host_scene("bunny", 1024, 1024, 4, 32768)
    gs = _gpu_scene(hs)
    px, st = gs.render()
    assert st.cameraSamples == 1024 * 1024 * 4
    w = px["weightSum"]
    assert w.min() >= 4 and w.max() <= 4 + 8 and abs(w.sum() - (4 * 1024 * 1024 + (w - 4).sum())) < 1
    assert np.isfinite(px["Lxyz"]).all() and (px["Lxyz"] >= 0).all()
    rd = hs.render_desc()
    rd.shardCount, rd.shardIndex, rd.shardBlock = 2, 1, 4096
    half, _ = gs.render(rd)
    own = half["weightSum"] >= 4
    assert abs(own.mean() - 0.5) < 0.01
    assert np.array_equal(half["Lxyz"][own & (w == 4)].view(np.uint32), px["Lxyz"][own & (w == 4)].view(np.uint32))


def test_synthetic_scene_bit_exact(staged):
    """BASELINE.json configs[4] generator (random small triangles, disk + point light, maxdepth 8) at a CI size."""
    hs = pb.HostScene.synthetic(30000, 48, 48, 8, maxdepth=8, ncores=72)
    px_gpu, st = _gpu_scene(hs).render()
    px_cpu, _ = oa.render(hs, threads=2)
    assert np.array_equal(px_gpu["weightSum"], px_cpu["weightSum"])
    diff = (px_gpu["Lxyz"].view(np.uint32) != px_cpu["Lxyz"].view(np.uint32)).any(axis=2)
    assert (diff & (px_cpu["weightSum"] == 8)).sum() <= 1
    assert st.cameraSamples == 48 * 48 * 8
