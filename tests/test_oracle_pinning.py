"""Pins the oracle (oracle/liboracle.so, the CPU restatement) against the reference:
(a) committed golden vectors minted from the UNMODIFIED reference (tests/golden/,
    generator: tests/golden/make_golden.py), and
(b) where oracle/_ref/bin/pbrt exists, the reference binary run on the spot.
Bit-exact except for the film pixels that receive samples from two different
render tasks (edge splats, film/image.cpp:82-89,122-127), where the reference itself
is order-dependent run to run."""
import json
import os

import numpy as np
import pytest

import pbrt_v2_b200 as pb
import oracle_api as oa
from helpers import host_scene

GOLD = os.path.join(pb.REPO, "tests", "golden")
pytestmark = pytest.mark.skipif(not oa.have_oracle(), reason="oracle/liboracle.so not built")


def test_rng_matches_reference_mt19937():
    g = np.load(os.path.join(GOLD, "rng.npz"))
    # the oracle exposes its RNG through the LD tables; check the MT19937 core via numpy's generator too
    for key in g.files:
        seed = int(key[4:])
        want = g[key]
        mt = np.random.RandomState()
        # pbrt seeds with init_genrand(seed) == numpy's legacy seeding for 32-bit ints
        mt.seed(seed)
        got = mt.randint(0, 2**32, size=len(want), dtype=np.uint64).astype(np.uint32)
        assert np.array_equal(got, want), "reference RNG is not plain MT19937 for seed %d" % seed


def test_ld_tables_match_golden():
    g = np.load(os.path.join(GOLD, "ld_samples.npz"))
    cases = [k for k in g.files if not k.endswith("_next")]
    assert len(cases) >= 5
    for key in cases:
        res, spp, ntasks, task = [int(x[1:]) for x in key.split("_")]
        rd = pb.set_path_layout(pb.RenderDesc())
        rd.xResolution = rd.yResolution = res
        rd.cropWindow[:] = [0, 1, 0, 1]
        rd.filterXWidth = rd.filterYWidth = 0.5
        rd.samplesPerPixel = spp
        rd.maxDepth = 5
        rd.nTasks, rd.taskBegin, rd.taskEnd = ntasks, 0, ntasks
        rd.shutterOpen, rd.shutterClose = 0.0, 1.0
        got = oa.ld_samples(rd, task)
        want = g[key]
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), key


def test_oracle_ld_samples_general_layouts_match_golden():
    """LDPixelSample with arrays of more than one value per pixel sample (DirectLightingIntegrator's layouts),
    against vectors minted from the reference's own Sample / LDSampler classes (refdump ldx)."""
    from helpers import layout_rd
    g = np.load(os.path.join(GOLD, "ld_samples_general.npz"))
    cases = [k for k in g.files if not k.endswith("_next") and not k.endswith("_meta")]
    assert len(cases) >= 4
    for key in cases:
        rd, task = layout_rd(g[key + "_meta"])
        got = oa.ld_samples(rd, task)
        assert got.shape == g[key].shape
        assert np.array_equal(got.view(np.uint32), g[key].view(np.uint32)), key


GOLDEN_IMAGES = ["bunny_r32_s4_c32", "bunny_r32_s8_c1", "killeroo_simple_r32_s4_c32", "killeroo_simple_r32_s8_c1",
                 "anim_killeroos_moving_r32_s4_c32", "anim_killeroos_moving_r32_s8_c1",
                 "directlighting_x_killeroo_simple_r32_s4_c32", "directlighting_x_killeroo_simple_r32_s8_c1",
                 "directlighting_x_bunny_r32_s4_c32", "directlightingone_x_bunny_r32_s8_c1",
                 "directlighting_x_anim_killeroos_moving_r32_s4_c1"]


@pytest.mark.parametrize("key", GOLDEN_IMAGES)
def test_oracle_image_matches_golden_reference_render(staged, key):
    g = np.load(os.path.join(GOLD, "images.npz"))
    from helpers import golden_image_case
    base, res, spp, ncores, integrator, strategy = golden_image_case(key)
    hs = host_scene(base, res, res, spp, ncores, integrator=integrator, strategy=strategy)
    px, _ = oa.render(hs, threads=4)
    rgb = hs.pixels_to_rgb(px)
    want = g[key]
    differ = (rgb.view(np.uint32) != want.view(np.uint32)).any(axis=2)
    # multi-pixel tasks (c1): every pixel bit-exact; per-pixel tasks: only cross-task edge splats may differ
    assert differ.sum() <= (0 if ncores <= 1 else 4), "%d pixels differ from the reference render" % differ.sum()
    np.testing.assert_allclose(rgb, want, rtol=2e-6, atol=0)


@pytest.mark.parametrize("base", ["bunny", "killeroo-simple"])
def test_aggregatetest_matches_reference_output(staged, base):
    """The reference's accelerator check (Renderer "aggregatetest", renderers/aggregatetest.cpp:60-119), restated: the
    same 20 000 rays (RNG feedback through the last hit point / epsilon included), BVHAccel walk vs the exhaustive loop.
    Known answer from the unmodified binary (tests/golden/aggregatetest.json): the number of "Disagreement" warnings
    (rays leaving a flat primitive with mint == 0, which BVHAccel's strict slab test misses) and the rays it printed."""
    g = json.load(open(os.path.join(GOLD, "aggregatetest.json")))[base]
    hs = host_scene(base, 32, 32, 4, 32)
    bad, rays, hit, t = oa.aggregate_test(hs, g["niters"])
    assert bad == g["disagreements"]
    mine = {(tuple(rays[i]["o"].tolist()), tuple(rays[i]["d"].tolist())) for i in range(len(rays))}
    for r in g["first_rays"]:
        key = (tuple(np.float32(r["o"]).tolist()), tuple(np.float32(r["d"]).tolist()))
        assert key in mine
    assert 0.05 < hit.mean() < 0.9
    # the batch entry point walks the BVH exactly like the in-test accelerator call
    again = oa.intersect(hs, rays)
    walk_disagrees = (again["hit"].astype(np.uint8) != hit) | (again["t"].view(np.uint32) != t.view(np.uint32))
    assert walk_disagrees.sum() >= bad   # (= bad + the "inconsistent bounds" cases the reference does not report)


def test_bxdf_restatements_match_reference_classes():
    """Lambertian, Microfacet(Blinn) f / Pdf / Sample_f and FresnelDielectric(1.5, 1) against values printed by the
    reference's own classes (refdump bxdf; tools/bsdftest.cpp exercises the same ones), bit for bit, incl. powf."""
    g = np.load(os.path.join(GOLD, "bxdf.npz"))
    for key in [k for k in g.files if not k.endswith("_exponent")]:
        rows = g[key]
        got = oa.bxdf_test(float(g[key + "_exponent"]), rows[:, 0:3], rows[:, 3:6], rows[:, 6:8])
        want = rows[:, 8:22]
        for c, name in enumerate(["lambert f", "lambert pdf", "blinn f", "blinn pdf", "blinn wi.x", "blinn wi.y", "blinn wi.z",
                                  "blinn sample pdf", "blinn sample f", "lambert wi.x", "lambert wi.y", "lambert wi.z",
                                  "lambert sample pdf", "fresnel"]):
            same = got[:, c].view(np.uint32) == want[:, c].view(np.uint32)
            both_nan = np.isnan(got[:, c]) & np.isnan(want[:, c])
            assert (same | both_nan).all(), "%s: %s differs on %d of %d" % (key, name, (~(same | both_nan)).sum(), len(same))


def test_exrdiff_metric_on_golden_images(staged):
    """tools/exrdiff.cpp's thresholds (the reference's definition of "same image") on oracle vs reference renders"""
    from helpers import exrdiff, golden_image_case
    g = np.load(os.path.join(GOLD, "images.npz"))
    for key in GOLDEN_IMAGES:
        base, res, spp, ncores, integrator, strategy = golden_image_case(key)
        hs = host_scene(base, res, res, spp, ncores, integrator=integrator, strategy=strategy)
        px, _ = oa.render(hs, threads=4)
        small, big, avg = exrdiff(g[key], hs.pixels_to_rgb(px))
        assert (small, big) == (0, 0) and abs(avg) < 1e-6, key


def test_bvh_sizes_match_reference(staged):
    counts = json.load(open(os.path.join(GOLD, "counts.json")))
    for base, want in counts.items():
        got = host_scene(base, 8, 8, 1, 1).counts()
        assert got["nodes"] == want["nodes"] and got["primitives"] == want["primitives"], (base, got, want)


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("base,res,spp,ncores,integrator", [
    ("bunny", 40, 8, 2, "path"), ("killeroo-simple", 40, 4, 2, "path"), ("anim-killeroos-moving", 40, 4, 2, "path"),
    ("killeroo-simple", 40, 4, 2, "directlighting"), ("bunny", 40, 4, 2, "directlighting")])
def test_oracle_bit_exact_vs_reference_binary(staged, base, res, spp, ncores, integrator):
    hs = host_scene(base, res, res, spp, ncores, tag="-live", integrator=integrator)
    px, _ = oa.render(hs, threads=4)
    rgb = hs.pixels_to_rgb(px)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(rgb.view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["dof", "maxdepth1", "maxdepth8", "nonsquare-crop", "directlighting-maxdepth1", "shutter",
                                  "spp1", "film1x1", "maxdepth0", "nolights", "nolights-directlighting"])
def test_oracle_bit_exact_vs_reference_binary_on_variants(staged, name):
    """thin-lens camera (perspective.cpp:100-113), maxdepth 1 / 8, a non-square cropped film, directlighting without the
    SpecularReflect/Transmit draws (maxdepth 1), a shutter interval inside the transform interval (animated scene)"""
    from helpers import variant_scene
    hs, ncores = variant_scene(name)
    px, _ = oa.render(hs, threads=4)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["path-spectrum-params", "direct-spectrum-light", "path-blackbody-light"])
def test_spectrum_parameters_match_reference_binary(staged, name):
    """"spectrum" parameters from .spd files (gold's measured eta and k for MetalMaterial, the D65 illuminant for a point
    light) and from inline wavelength / value pairs given out of order, "xyz" parameters, "blackbody" parameters (Planck's
    law over the CIE wavelengths, spectrum.cpp:195-204): RGBSpectrum::FromSampled / FromXYZ as the reference's ParamSet applies
    them (paramset.cpp:109-187, spectrum.h:428-451)"""
    from helpers import variant_scene
    hs, ncores = variant_scene(name)
    px, _ = oa.render(hs, threads=4)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["path-heightfield", "direct-heightfield", "path-discard-degenerate-uvs", "path-keep-degenerate-uvs"])
def test_heightfield_shape_matches_reference_binary(staged, name):
    """Shape "heightfield" (shapes/heightfield.cpp:69-130): the nu x nv grid refined into one triangle mesh, two triangles per
    cell in the reference's winding, uv = xy; under a transform with rotation and non-uniform use of the unit square.  And
    "bool discarddegenerateUVs" on a triangle mesh with normals (trianglemesh.cpp:402-421): on and off"""
    from helpers import variant_scene
    hs, ncores = variant_scene(name)
    px, _ = oa.render(hs, threads=4)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["path-constant-textures", "direct-constant-textures"])
def test_constant_textures_match_reference_binary(staged, name):
    """Texture "constant" / "scale" / "mix" (textures/{constant,scale,mix}.h: values that do not depend on the surface point,
    folded into the material by the front end) given to Kd, Ks and roughness through "texture" parameters, nested; an undefined
    texture name (Error, then the parameter's constant, core/paramset.cpp:598-630)"""
    from helpers import variant_scene
    hs, ncores = variant_scene(name)
    px, _ = oa.render(hs, threads=4)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["path-instances-bunny", "direct-instances-killeroo"])
def test_object_instances_match_reference_binary(staged, name):
    """ObjectBegin / ObjectEnd / ObjectInstance (core/api.cpp:1095-1147): every instance is a TransformedPrimitive over the
    object's aggregate (built once with the scene's Accelerator) or over its single sphere; static
    and moving instance transforms, non-uniform scale, mirror instances, unused and empty objects"""
    from helpers import variant_scene
    hs, ncores = variant_scene(name)
    px, _ = oa.render(hs, threads=4)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["path-named-materials", "direct-named-materials"])
def test_named_materials_match_reference_binary(staged, name):
    """MakeNamedMaterial / NamedMaterial (core/api.cpp:939-968, 1062-1077): own parameters over the current Material's, an
    unknown name falling back to the current Material, Material resetting the selection, AttributeEnd restoring it"""
    from helpers import variant_scene
    hs, ncores = variant_scene(name)
    px, _ = oa.render(hs, threads=4)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["direct-infinite", "path-infinite-bunny", "direct-envmap-npot"])
def test_quick_render_option_matches_reference_binary(staged, name):
    """--quick (PbrtOptions.quickRender): film resolution / 4 (film/image.cpp:244-245), one pixel sample
    (lowdiscrepancy.cpp:88), light samples / 4 (diffuse.cpp:64, infinite.cpp:190)"""
    from helpers import variant_scene
    hs0, ncores = variant_scene(name)
    hs = pb.HostScene(hs0.filename, ncores=ncores, quick=True)
    rd = hs.render_desc()
    assert (rd.xResolution, rd.yResolution, rd.samplesPerPixel) == (10, 10, 1)
    px, _ = oa.render(hs, threads=4)
    ref, _ = oa.run_reference(hs.filename, ncores, extra_args=["--quick"])
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["path-glass-mirror", "path-glass-params-mirror-floor", "direct-glass-mirror",
                                  "direct-mirror-floor-maxdepth3", "direct-one-glass"])
def test_oracle_specular_materials_bit_exact_vs_reference_binary(staged, name):
    """mirror / glass (SpecularReflection, SpecularTransmission; materials/mirror.cpp, glass.cpp): the specularBounce
    branches of PathIntegrator (path.cpp:63,98) and the recursive SpecularReflect / SpecularTransmit rays of
    DirectLightingIntegrator (integrator.cpp:177-258)"""
    from helpers import variant_scene
    hs, ncores = variant_scene(name)
    px, _ = oa.render(hs, threads=4)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["path-metal-copper", "path-metal-params-floor", "direct-metal"])
def test_oracle_metal_material_bit_exact_vs_reference_binary(staged, name):
    """MetalMaterial (materials/metal.cpp): Microfacet(1, FresnelConductor(eta, k), Blinn(1 / roughness)) -- the default
    copper coefficients (RGBSpectrum::FromSampled of the measured n and k, read out of the reference by `refdump metal`),
    rgb-valued eta / k, both integrators"""
    from helpers import variant_scene
    hs, ncores = variant_scene(name)
    px, _ = oa.render(hs, threads=4)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["path-infinite", "path-infinite-specular", "direct-infinite", "direct-one-infinite",
                                  "path-infinite-bunny"])
def test_oracle_infinite_light_bit_exact_vs_reference_binary(staged, name):
    """InfiniteAreaLight with a constant radiance (lights/infinite.cpp over its 1x1 map: MIPMap::Lookup's bilinear weights,
    Distribution2D sampling, Pdf): Le of escaping camera rays (samplerrenderer.cpp:238-242), of escaping specular bounces
    (path.cpp:113-117, and through SpecularReflect / SpecularTransmit under directlighting), of escaping MIS rays
    (integrator.cpp:163-164), the light sample itself; rotated light-to-world transform, nsamples, both strategies"""
    from helpers import variant_scene
    hs, ncores = variant_scene(name)
    px, _ = oa.render(hs, threads=4)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["path-envmap-pow2", "path-envmap-npot-specular", "direct-envmap-npot", "direct-one-envmap-grey",
                                  "path-envmap-bunny", "path-envmap-1x1", "path-envmap-tga", "direct-envmap-tga-mapped"])
def test_oracle_environment_maps_bit_exact_vs_reference_binary(staged, name):
    """InfiniteAreaLight with "mapname" (.pfm): the host front end restates the constructor (ReadImagePFM, texels x L x scale,
    MIPMap's Lanczos resampling to a power of two and its pyramid, the filtered luminance x sin(theta) image, Distribution2D
    -- pbrt_v2_b200/csrc/host/envmap.cpp), the oracle the per-ray work (bilinear MIPMap::Lookup over level 0 with
    TEXTURE_REPEAT, Distribution2D::SampleContinuous / Pdf).  Power-of-two and other resolutions, RGB and grey files, both byte
    orders and a scale word, a 1x1 file, L and scale on top; both integrators and strategies."""
    from helpers import variant_scene
    hs, ncores = variant_scene(name)
    px, _ = oa.render(hs, threads=4)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
@pytest.mark.parametrize("name", ["filter-gaussian", "filter-mitchell", "filter-triangle-crop", "filter-sinc", "filter-widebox", "filter-narrowbox"])
def test_oracle_wide_pixel_filters_bit_exact_vs_reference_binary(staged, name):
    """gaussian / mitchell / triangle / sinc / wide box filters (filters/*.cpp; ImageFilm's 16 x 16 weight table and the
    sample's raster extent, film/image.cpp:57-69, 77-112; the sample extent grows with the filter, :157-166): every pixel
    of the image, bit for bit, at --ncores 1 -- EnqueueTasks then runs the tasks inline in enqueue order, nTasks-1 down to 0
    (parallel.cpp:800-805), so the reference's own float sums are ordered; the oracle adds in the same order with one thread."""
    from helpers import variant_scene
    hs, ncores = variant_scene(name)
    assert ncores == 1
    px, _ = oa.render(hs, threads=1)
    ref, _ = oa.run_reference(hs.filename, ncores)
    assert np.array_equal(hs.pixels_to_rgb(px).view(np.uint32), ref.view(np.uint32))


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
def test_synthetic_scene_text_and_direct_routes_match_reference(staged):
    """the C5 generator: the in-memory build equals the text scene, and the oracle renders it bit-exactly
    like the unmodified reference (multi-pixel tasks)."""
    from pbrt_v2_b200 import scenes
    path = pb.HostScene.write_synthetic(os.path.join(scenes.STAGED, "t-synth.pbrt"), 5000, 32, 32, 4, 8, "t-synth.pfm")
    direct = pb.HostScene.synthetic(5000, 32, 32, 4, maxdepth=8, ncores=1)
    text = pb.HostScene(path, ncores=1)
    assert direct.counts() == text.counts() and direct.scene_bytes() == text.scene_bytes()
    a, _ = oa.render(direct, threads=2)
    b, _ = oa.render(text, threads=2)
    assert np.array_equal(a["Lxyz"].view(np.uint32), b["Lxyz"].view(np.uint32))
    ref, _ = oa.run_reference(path, 1)
    assert np.array_equal(direct.pixels_to_rgb(a).view(np.uint32), ref.view(np.uint32))


def _read_png(path):
    """8-bit RGB PNG -> uint8 [h, w, 3] (all five scan-line filters; what stb_image_write and zlib-based writers produce)"""
    import struct, zlib
    d = open(path, "rb").read()
    assert d[:8] == b"\x89PNG\r\n\x1a\n"
    p, idat, w, h = 8, b"", 0, 0
    while p < len(d):
        n, typ = struct.unpack(">I4s", d[p:p + 8])
        body = d[p + 8:p + 8 + n]
        assert zlib.crc32(typ + body) == struct.unpack(">I", d[p + 8 + n:p + 12 + n])[0]
        if typ == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 2)
        elif typ == b"IDAT":
            idat += body
        p += 12 + n
    raw = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + 3 * w)
    out = np.zeros((h, 3 * w), np.int32)
    for y in range(h):
        f, line = int(raw[y, 0]), raw[y, 1:].astype(np.int32)
        up = out[y - 1] if y else np.zeros(3 * w, np.int32)
        cur = out[y]
        for i in range(3 * w):
            a = cur[i - 3] if i >= 3 else 0
            b, c = up[i], (up[i - 3] if i >= 3 else 0)
            if f == 0: pred = 0
            elif f == 1: pred = a
            elif f == 2: pred = b
            elif f == 3: pred = (a + b) // 2
            else:
                pa, pb_, pc = abs(b - c), abs(a - c), abs(a + b - 2 * c)
                pred = a if pa <= pb_ and pa <= pc else (b if pb_ <= pc else c)
            cur[i] = (line[i] + pred) & 255
    return out.reshape(h, w, 3).astype(np.uint8)


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
def test_png_film_output_matches_reference_binary(staged, tmp_path):
    """ImageFilm::WriteImage to a .png name (core/imageio.cpp:107-126: 8 bits, gamma 2.2, blue in the first channel as the
    reference stores it): the decoded image of the front end's file = the decoded image of the reference's own"""
    from pbrt_v2_b200 import scenes
    path = scenes.variant("bunny", os.path.join(scenes.STAGED, "t-png.pbrt"), 24, 20, 4, filename="t-png.png")
    ref_png = os.path.join(scenes.STAGED, "t-png.png")
    if os.path.exists(ref_png):
        os.remove(ref_png)
    import subprocess
    subprocess.run([oa.REF_BIN, "--ncores", "2", "--quiet", os.path.basename(path)], cwd=scenes.STAGED, check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    want = _read_png(ref_png)
    hs = pb.HostScene(path, ncores=2)
    px, _ = oa.render(hs, threads=2)
    mine = str(tmp_path / "mine.png")
    hs.write_image(px, mine)
    got = _read_png(mine)
    assert got.shape == want.shape == (20, 24, 3)
    assert np.array_equal(got, want)
    assert int(got.max()) > 100 and not np.array_equal(got[..., 0], got[..., 2])


@pytest.mark.skipif(not oa.have_reference(), reason="oracle/_ref/bin/pbrt not built")
def test_tga_and_pfm_film_output_are_the_reference_binarys_files(staged, tmp_path):
    """ImageFilm::WriteImage to a .tga name (WriteImageTGA, core/imageio.cpp:200-225, over targa.c's tga_write_bgr: uncompressed
    24-bit BGR, gamma 2.2, top-to-bottom rows, the TGA 2.0 footer) and to a .pfm name (WriteImagePFM, :410-455): the files, byte
    for byte; a crop window gives the cropped image"""
    import subprocess
    from pbrt_v2_b200 import scenes
    for tag, crop, ext in (("full", None, "tga"), ("crop", (0.25, 0.75, 0.1, 0.6), "tga"), ("pfm", (0.25, 0.75, 0.1, 0.6), "pfm")):
        path = scenes.variant("bunny", os.path.join(scenes.STAGED, "t-tga-%s.pbrt" % tag), 24, 20, 4, filename="t-tga-%s.%s" % (tag, ext), crop=crop)
        ref_tga = os.path.join(scenes.STAGED, "t-tga-%s.%s" % (tag, ext))
        if os.path.exists(ref_tga):
            os.remove(ref_tga)
        subprocess.run([oa.REF_BIN, "--ncores", "1", "--quiet", os.path.basename(path)], cwd=scenes.STAGED, check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        hs = pb.HostScene(path, ncores=1)
        px, _ = oa.render(hs, threads=1)
        mine = str(tmp_path / ("mine-%s.%s" % (tag, ext)))
        hs.write_image(px, mine)
        assert open(mine, "rb").read() == open(ref_tga, "rb").read(), tag
