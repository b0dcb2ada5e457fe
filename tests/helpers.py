"""Shared helpers for the parity tests."""
import os
import numpy as np

import pbrt_v2_b200 as pb
from pbrt_v2_b200 import scenes

_cache = {}


def host_scene(base, xres, yres, spp, ncores, maxdepth=None, crop=None, tag="", nlevels=None, integrator="path",
               strategy=None):
    key = (base, xres, yres, spp, ncores, maxdepth, crop, tag, nlevels, integrator, strategy)
    if key not in _cache:
        name = "%s-%dx%d-%dspp%s%s%s.pbrt" % (base, xres, yres, spp, tag, "" if nlevels is None else "-nl%d" % nlevels,
                                              "" if integrator == "path" else "-%s%s" % (integrator, strategy or ""))
        path = scenes.variant(base, scenes.STAGED + "/" + name, xres, yres, spp, maxdepth=maxdepth, crop=crop,
                              nlevels=nlevels, integrator=integrator, strategy=strategy)
        _cache[key] = pb.HostScene(path, ncores=ncores)
    return _cache[key]


def random_rays(n, seed, bounds_lo, bounds_hi, inside_frac=0.5):
    """aggregatetest-style rays (renderers/aggregatetest.cpp:61-119): random origins in/around
    the scene bounds, random directions, a share of axis-aligned directions."""
    rng = np.random.RandomState(seed)
    lo, hi = np.asarray(bounds_lo, np.float32), np.asarray(bounds_hi, np.float32)
    ext = hi - lo
    rays = np.zeros(n, dtype=pb.RAY_DTYPE)
    o = lo - ext + rng.rand(n, 3).astype(np.float32) * 3 * ext
    inside = rng.rand(n) < inside_frac
    o[inside] = lo + rng.rand(inside.sum(), 3).astype(np.float32) * ext
    target = lo + rng.rand(n, 3).astype(np.float32) * ext
    d = target - o
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    axis = rng.rand(n) < 0.05
    k = rng.randint(0, 3, n)
    da = np.zeros((n, 3), np.float32)
    da[np.arange(n), k] = np.where(rng.rand(n) < 0.5, 1.0, -1.0)
    d[axis] = da[axis]
    rays["o"] = o
    rays["d"] = d.astype(np.float32)
    rays["mint"] = 0.0
    rays["maxt"] = np.inf
    short = rng.rand(n) < 0.2
    rays["maxt"][short] = (rng.rand(short.sum()) * np.linalg.norm(ext)).astype(np.float32)
    return rays


def scene_bounds(hs):
    """root node bounds of the flattened BVH (first 6 floats of nodes[0])."""
    import ctypes as C
    desc = C.cast(hs.desc_ptr, C.POINTER(C.c_uint32 * 4))   # nNodes, pad, ptr...
    # PbrtGpuSceneDesc starts with {uint32 nNodes; (pad); const PbrtGpuBVHNode* nodes}
    nodes_ptr = C.cast(hs.desc_ptr + 8, C.POINTER(C.c_void_p))[0]
    b = np.ctypeslib.as_array(C.cast(nodes_ptr, C.POINTER(C.c_float)), shape=(6,)).copy()
    return b[:3], b[3:]


def golden_image_case(key):
    """(base, res, spp, ncores, integrator, strategy) of a tests/golden/images.npz key"""
    integrator, strategy = "path", None
    if "_x_" in key:
        head, key = key.split("_x_")
        integrator = "directlighting"
        strategy = head[len("directlighting"):] or None
    parts = key.split("_")
    res, spp, ncores = [int(x[1:]) for x in parts[-3:]]
    return "-".join(parts[:-3]), res, spp, ncores, integrator, strategy


def layout_rd(meta):
    """RenderDesc of a tests/golden/ld_samples_general.npz case (its `_meta` vector)"""
    m = [int(x) for x in meta]
    res, spp, ntasks, task, c1 = m[:5]
    s1 = m[5:5 + c1]
    c2 = m[5 + c1]
    s2 = m[6 + c1:6 + c1 + c2]
    rd = pb.RenderDesc()
    rd.xResolution = rd.yResolution = res
    rd.cropWindow[:] = [0, 1, 0, 1]
    rd.filterXWidth = rd.filterYWidth = 0.5
    rd.samplesPerPixel = spp
    rd.maxDepth = 5
    rd.integrator = 1
    rd.nTasks, rd.taskBegin, rd.taskEnd = ntasks, 0, ntasks
    rd.shutterOpen, rd.shutterClose = 0.0, 1.0
    rd.n1D, rd.n2D = c1, c2
    for i, v in enumerate(s1):
        rd.size1D[i] = v
    for i, v in enumerate(s2):
        rd.size2D[i] = v
    return rd, task


# scene-file variants beyond resolution / spp / integrator: (name, base, xres, yres, spp, ncores, variant kwargs, text edits)
VARIANTS = [
    ("dof", "bunny", 40, 40, 4, 2, {},
     [('Camera "perspective"', 'Camera "perspective" "float lensradius" [0.02] "float focaldistance" [0.6]')]),
    ("maxdepth1", "killeroo-simple", 40, 40, 4, 2, {"maxdepth": 1}, []),
    ("maxdepth8", "killeroo-simple", 40, 40, 4, 2, {"maxdepth": 8}, []),
    ("nonsquare-crop", "bunny", 48, 32, 4, 3, {"crop": (0.1, 0.9, 0.25, 0.8)}, []),
    ("directlighting-maxdepth1", "killeroo-simple", 40, 40, 4, 2, {"maxdepth": 1, "integrator": "directlighting"}, []),
    ("shutter", "anim-killeroos-moving", 40, 40, 4, 2, {},
     [('Camera "perspective"', 'Camera "perspective" "float shutteropen" [0.25] "float shutterclose" [0.5]')]),
]
# edge cases: one sample per pixel, a 1x1 film, maxdepth 0, scenes without lights under both integrators
EDGE_VARIANTS = [
    ("spp1", "bunny", 24, 24, 1, 1, {}, []),
    ("film1x1", "killeroo-simple", 1, 1, 4, 1, {}, []),
    ("maxdepth0", "killeroo-simple", 24, 24, 4, 2, {"maxdepth": 0}, []),
    ("nolights", "bunny", 24, 24, 4, 2, {},
     [('LightSource "point" "color I" [ .7 .7 .7 ]', ''), ('AreaLightSource "area" "color L" [15 15 15 ] "integer nsamples"  [4]', '')]),
    ("nolights-directlighting", "bunny", 24, 24, 4, 2, {"integrator": "directlighting"},
     [('LightSource "point" "color I" [ .7 .7 .7 ]', ''), ('AreaLightSource "area" "color L" [15 15 15 ] "integer nsamples"  [4]', '')]),
]


_M1 = 'Material "plastic" "color Kd" [.4 .2 .2] "color Ks" [.5 .5 .5]\n        "float roughness" [.025]'
_M2 = 'Material "plastic" "color Ks" [.3 .3 .3] "color Kd" [.4 .5 .4]\n        "float roughness" [.15] '
_FLOOR = 'Material "matte" "color Kd" [.5 .5 .8]'
# SURVEY 8(f) rank 4, specular materials: host front end + oracle so far (the device answers PBRTGPU_EUNSUPPORTED)
SPECULAR_VARIANTS = [
    ("path-glass-mirror", "killeroo-simple", 40, 40, 4, 2, {}, [(_M1, 'Material "glass"'), (_M2, 'Material "mirror"')]),
    ("path-glass-params-mirror-floor", "killeroo-simple", 40, 40, 4, 2, {"maxdepth": 8},
     [(_M1, 'Material "glass" "color Kr" [.8 .9 1] "color Kt" [.9 .7 .6] "float index" [1.33]'),
      (_FLOOR, 'Material "mirror" "color Kr" [.7 .7 .7]')]),
    ("direct-glass-mirror", "killeroo-simple", 40, 40, 4, 2, {"integrator": "directlighting"},
     [(_M1, 'Material "glass"'), (_M2, 'Material "mirror"')]),
    ("direct-mirror-floor-maxdepth3", "killeroo-simple", 40, 40, 4, 2, {"integrator": "directlighting", "maxdepth": 3},
     [(_FLOOR, 'Material "mirror"'), (_M1, 'Material "glass" "float index" [1.2]')]),
    ("direct-one-glass", "killeroo-simple", 40, 40, 4, 1, {"integrator": "directlighting", "strategy": "one"},
     [(_M1, 'Material "glass"')]),
]


# SURVEY 8(f) rank 4, MetalMaterial: host front end + oracle (pinned on the reference binary) + device code behind
# -DPBRT_METAL=1 (host-verified; the shipped library answers PBRTGPU_EUNSUPPORTED)
METAL_VARIANTS = [
    ("path-metal-copper", "killeroo-simple", 40, 40, 4, 2, {}, [(_M1, 'Material "metal"')]),
    ("path-metal-params-floor", "killeroo-simple", 40, 40, 4, 2, {"maxdepth": 8},
     [(_M2, 'Material "metal" "color eta" [.14 .37 1.44] "color k" [3.98 2.38 1.6] "float roughness" [.08]'),
      (_FLOOR, 'Material "metal" "float roughness" [.3]')]),
    ("direct-metal", "killeroo-simple", 40, 40, 4, 2, {"integrator": "directlighting"},
     [(_M1, 'Material "metal" "float roughness" [.05]'), (_M2, 'Material "metal"')]),
]


# SURVEY 8(f) rank 4, InfiniteAreaLight with a constant radiance: host front end + oracle (pinned on the reference binary); the
# device answers PBRTGPU_EUNSUPPORTED
_WB = "WorldBegin"
_INF1 = _WB + '\nAttributeBegin\nRotate 35 1 .2 0\nLightSource "infinite" "color L" [.4 .5 .7]\nAttributeEnd\n'
_INF2 = _WB + '\nLightSource "infinite" "color L" [.2 .2 .2] "color scale" [2 1 .5] "integer nsamples" [4]\n'
INFINITE_VARIANTS = [
    ("path-infinite", "killeroo-simple", 40, 40, 4, 2, {}, [(_WB, _INF1)]),
    ("path-infinite-specular", "killeroo-simple", 40, 40, 4, 2, {"maxdepth": 8},
     [(_WB, _INF2), (_M1, 'Material "glass"'), (_M2, 'Material "mirror"')]),
    ("direct-infinite", "killeroo-simple", 40, 40, 4, 2, {"integrator": "directlighting"}, [(_WB, _INF2), (_M2, 'Material "mirror"')]),
    ("direct-one-infinite", "bunny", 40, 40, 4, 1, {"integrator": "directlighting", "strategy": "one"}, [(_WB, _INF1)]),
    ("path-infinite-bunny", "bunny", 40, 40, 8, 64, {}, [(_WB, _INF1)]),
]

# ... with a radiance map ("mapname"): .pfm files written next to the staged scene by variant_scene (ENV_MAPS: file name ->
# width, height, seed, channels, PFM scale word: negative = little-endian data, |scale| multiplies).  Powers of two and not
# (MIPMap's Lanczos resampling; the distribution keeps the file's resolution), RGB and grey, both byte orders, 1x1.
ENV_MAPS = {
    "env-16x8.pfm": (16, 8, 1, 3, -1.0),
    "env-12x7.pfm": (12, 7, 2, 3, -1.0),
    "env-grey-8x8-be.pfm": (8, 8, 3, 1, 2.0),
    "env-32x16.pfm": (32, 16, 4, 3, -0.5),
    "env-1x1.pfm": (1, 1, 5, 3, -1.0),
    "env-5x3.pfm": (5, 3, 6, 3, -1.0),
}
def _inf_map(fn, extra=""):
    return _WB + '\nAttributeBegin\nRotate 35 1 .2 0\nLightSource "infinite" "string mapname" "%s"%s\nAttributeEnd\n' % (fn, extra)
ENVMAP_VARIANTS = [
    ("path-envmap-pow2", "killeroo-simple", 40, 40, 4, 2, {}, [(_WB, _inf_map("env-16x8.pfm"))]),
    ("path-envmap-npot-specular", "killeroo-simple", 40, 40, 4, 2, {"maxdepth": 8},
     [(_WB, _inf_map("env-12x7.pfm", ' "color L" [.5 .6 .9] "integer nsamples" [2]')), (_M1, 'Material "glass"'), (_M2, 'Material "mirror"')]),
    ("direct-envmap-npot", "killeroo-simple", 40, 40, 4, 2, {"integrator": "directlighting"},
     [(_WB, _inf_map("env-5x3.pfm", ' "integer nsamples" [4]')), (_M2, 'Material "mirror"')]),
    ("direct-one-envmap-grey", "bunny", 40, 40, 4, 1, {"integrator": "directlighting", "strategy": "one"},
     [(_WB, _inf_map("env-grey-8x8-be.pfm", ' "color scale" [.2 .3 .25]'))]),
    ("path-envmap-bunny", "bunny", 40, 40, 8, 64, {}, [(_WB, _inf_map("env-32x16.pfm"))]),
    ("path-envmap-1x1", "bunny", 32, 32, 4, 2, {}, [(_WB, _inf_map("env-1x1.pfm"))]),
    # Targa files (ReadImageTGA over targa.c: restated in csrc/host/tga.cpp)
    ("path-envmap-tga", "bunny", 32, 32, 4, 2, {}, [(_WB, _inf_map("env-bgra-rle-td-16x8.tga", ' "color L" [3 3 3]'))]),
    ("direct-envmap-tga-mapped", "killeroo-simple", 32, 32, 4, 2, {"integrator": "directlighting"},
     [(_WB, _inf_map("env-mapped-12x6.tga", ' "color L" [2 2 2] "integer nsamples" [2]'))]),
]


def write_env_pfm(path, w, h, seed, channels, scale):
    """a small HDR-ish radiance map: log-normal texels, one very bright one, a black row"""
    rng = np.random.RandomState(seed)
    a = np.exp(rng.randn(h, w, channels)).astype(np.float32) * 0.3
    a[rng.randint(h), rng.randint(w)] = 25.0
    if h > 2:
        a[h - 1] = 0.0
    with open(path, "wb") as f:
        f.write(("PF" if channels == 3 else "Pf").encode() + b"\n%d %d\n%g\n" % (w, h, scale))
        f.write(a.astype("<f4" if scale < 0 else ">f4").tobytes())


def write_tga(path, w, h, seed, kind="bgr", rle=False, top_down=False, right_left=False, alpha=False, map_origin=0):
    """a Targa file of one of the kinds ReadImageTGA interprets: "bgr" (24 / 32 bit with alpha), "mono" (8 bit),
    "mapped" (8-bit indices into a 24 / 32-bit colour map whose first entry is index map_origin); raw or RLE packets;
    any scan direction.  Long constant runs make the RLE packets real."""
    rng = np.random.RandomState(seed)
    bpp = {"bgr": 4 if alpha else 3, "mono": 1, "mapped": 1}[kind]
    cmap = b""
    if kind == "mapped":
        n = 37
        mbpp = 4 if alpha else 3
        cmap = rng.randint(0, 256, (n, mbpp)).astype(np.uint8).tobytes()
        px = rng.randint(map_origin, map_origin + n, (h, w, 1)).astype(np.uint8)
    else:
        px = rng.randint(0, 256, (h, w, bpp)).astype(np.uint8)
    px[h // 2, : max(1, w - 1)] = px[h // 2, 0]
    flat = px.reshape(-1, bpp)
    if rle:
        data, i, n = bytearray(), 0, len(flat)
        while i < n:
            j = i + 1
            while j < n and j - i < 128 and (flat[j] == flat[i]).all():
                j += 1
            if j - i > 1:
                data += bytes([0x80 | (j - i - 1)]) + flat[i].tobytes()
            else:
                j = i + 1
                while j < n and j - i < 128 and not (j + 1 < n and (flat[j] == flat[j + 1]).all()):
                    j += 1
                data += bytes([j - i - 1]) + flat[i:j].tobytes()
            i = j
        data = bytes(data)
    else:
        data = flat.tobytes()
    typ = {"bgr": 2, "mono": 3, "mapped": 1}[kind] + (8 if rle else 0)
    ident = b"id!"
    head = bytes([len(ident), 1 if kind == "mapped" else 0, typ])
    head += np.array([map_origin if kind == "mapped" else 0, 37 if kind == "mapped" else 0], "<u2").tobytes()
    head += bytes([(32 if alpha else 24) if kind == "mapped" else 0])
    head += np.array([0, 0, w, h], "<u2").tobytes() + bytes([8 * bpp, (0x20 if top_down else 0) | (0x10 if right_left else 0) | (8 if alpha and kind == "bgr" else 0)])
    with open(path, "wb") as f:
        f.write(head + ident + cmap + data)


TGA_MAPS = {   # file name -> write_tga arguments after the path
    "env-bgr-9x5.tga": (9, 5, 11, "bgr"),
    "env-bgra-rle-td-16x8.tga": (16, 8, 12, "bgr", True, True, False, True),
    "env-mono-rl-7x4.tga": (7, 4, 13, "mono", False, False, True),
    "env-mono-rle-8x8.tga": (8, 8, 14, "mono", True, True, True),
    "env-mapped-12x6.tga": (12, 6, 15, "mapped", False, False, False, False, 3),
    "env-mapped-rle-a-5x9.tga": (5, 9, 16, "mapped", True, True, False, True, 0),
}


# MakeNamedMaterial / NamedMaterial (core/api.cpp:939-968, 1062-1077): a named material made from its own parameters on top of
# the current Material's; selecting it; Material switching back; the attribute stack restoring the selection
_NM_DEF = _WB + ('\nMaterial "plastic" "color Ks" [.6 .1 .1]\nMakeNamedMaterial "shiny" "string type" "plastic" "color Kd" [.1 .4 .1] "float roughness" [.05]'
                 '\nMakeNamedMaterial "chalk" "string type" "matte" "color Kd" [.7 .7 .2]\nMaterial "matte" "color Kd" [.3 .3 .8]\n')
NAMED_MATERIAL_VARIANTS = [
    ("path-named-materials", "killeroo-simple", 40, 40, 4, 2, {},
     [(_WB, _NM_DEF), (_M1, 'NamedMaterial "shiny"'), (_M2, 'AttributeBegin\nNamedMaterial "chalk"\nAttributeEnd\nNamedMaterial "nosuchname"')]),
    ("direct-named-materials", "killeroo-simple", 40, 40, 4, 2, {"integrator": "directlighting"},
     [(_WB, _NM_DEF), (_M1, 'NamedMaterial "chalk"'), (_M2, 'NamedMaterial "shiny"\nMaterial "matte"')]),
]


# Texture (core/api.cpp:910-936) of the classes that are the same everywhere on a surface: "constant", "scale", "mix" (also of each
# other); float and colour; a material parameter given as a texture; a name that does not exist (Error + the constant default)
_TEX = _WB + ('\nTexture "green" "color" "constant" "color value" [.1 .6 .2]\nTexture "half" "float" "constant" "float value" [.5]\n'
              'Texture "dim" "color" "scale" "texture tex1" "green" "color tex2" [.9 .5 2.]\n'
              'Texture "blend" "color" "mix" "texture tex1" "dim" "color tex2" [.8 .8 .1] "float amount" [.3]\n'
              'Texture "rough" "float" "mix" "float tex1" [.02] "float tex2" [.3] "texture amount" "half"\n'
              'Texture "rougher" "float" "scale" "texture tex1" "rough" "float tex2" [1.5]\n')
TEXTURE_VARIANTS = [
    ("path-constant-textures", "killeroo-simple", 40, 40, 4, 2, {},
     [(_WB, _TEX), (_M1, 'Material "plastic" "texture Kd" "blend" "color Ks" [.5 .5 .5] "texture roughness" "rougher"'),
      (_M2, 'Material "matte" "texture Kd" "nosuchtexture" "color Kd" [.2 .3 .7]')]),
    ("direct-constant-textures", "killeroo-simple", 40, 40, 4, 2, {"integrator": "directlighting"},
     [(_WB, _TEX), (_M1, 'Material "plastic" "texture Kd" "dim" "texture Ks" "green" "texture roughness" "rough"'), (_M2, 'Material "matte" "texture Kd" "blend"')]),
]


# "spectrum" parameters (ParamSet::AddSampledSpectrum / AddSampledSpectrumFiles, paramset.cpp:138-187: .spd files and inline
# wavelength / value pairs -- unsorted ones too -- through RGBSpectrum::FromSampled) and "xyz" parameters
SPECTRUM_VARIANTS = [
    ("path-spectrum-params", "killeroo-simple", 40, 40, 4, 2, {},
     [(_M1, 'Material "metal" "spectrum eta" "spds/metals/Au.eta.spd" "spectrum k" "spds/metals/Au.k.spd" "float roughness" [.05]'),
      (_M2, 'Material "plastic" "spectrum Kd" [400 .2 500 .6 700 .4 600 .9] "xyz Ks" [.3 .35 .2] "float roughness" [.1]')]),
    ("path-blackbody-light", "bunny", 40, 40, 4, 2, {},
     [(_WB, _WB + '\nLightSource "point" "blackbody I" [3200 .08] "point from" [.3 .4 .3]\nLightSource "point" "blackbody I" [9000 .03] "point from" [-.3 .4 .2]\n')]),
    ("direct-spectrum-light", "bunny", 40, 40, 4, 2, {"integrator": "directlighting"},
     [(_WB, _WB + '\nLightSource "point" "spectrum I" "spds/cie.stdillum.D6500.spd" "color scale" [.05 .05 .05] "point from" [.3 .4 .3]\n')]),
]


# Shape "heightfield" (shapes/heightfield.cpp: refined into one triangle mesh with uv = xy)
def _heightfield(nu, nv, seed, place):
    z = np.random.RandomState(seed).uniform(0, .35, nu * nv)
    return (_WB + '\nAttributeBegin\n%s\nMaterial "plastic" "color Kd" [.3 .5 .7] "color Ks" [.3 .3 .3] "float roughness" [.1]\n'
            'Shape "heightfield" "integer nu" [%d] "integer nv" [%d] "float Pz" [%s]\nAttributeEnd\n' % (place, nu, nv, " ".join("%.6g" % v for v in z)))
# "bool discarddegenerateUVs" (trianglemesh.cpp:379, 402-421): with normals present, uvs that coincide on a triangle are dropped
# for the whole mesh -- which changes dpdu / dpdv, hence the shading frame, hence the image
_DEGEN = _WB + ('\nAttributeBegin\nTranslate -.12 .05 .1\nScale .25 .25 .25\nMaterial "plastic" "color Kd" [.6 .4 .2] "color Ks" [.4 .4 .4] "float roughness" [.05]\n'
                'Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [0 0 0 1 0 .2 1 1 0 0 1 .3] "normal N" [0 0 1 .2 0 1 0 .2 1 -.2 0 1]\n'
                '  "float uv" [0 0 1 0 1 0 0 1] "bool discarddegenerateUVs" ["%s"]\nAttributeEnd\n')
DEGENERATE_UV_VARIANTS = [
    ("path-discard-degenerate-uvs", "bunny", 40, 40, 4, 2, {}, [(_WB, _DEGEN % "true")]),
    ("path-keep-degenerate-uvs", "bunny", 40, 40, 4, 2, {}, [(_WB, _DEGEN % "false")]),
]
HEIGHTFIELD_VARIANTS = [
    ("path-heightfield", "bunny", 40, 40, 4, 2, {}, [(_WB, _heightfield(9, 7, 1, "Translate -.15 .02 -.1\nRotate -90 1 0 0\nScale .3 .3 .3"))]),
    ("direct-heightfield", "killeroo-simple", 40, 40, 4, 2, {"integrator": "directlighting"},
     [(_WB, _heightfield(5, 12, 2, "Translate -150 20 -100\nScale 300 300 300"))]),
]


# ObjectBegin / ObjectEnd / ObjectInstance (core/api.cpp:1095-1147): an object of several shapes (its own BVH, shared by the
# instances), an object of one sphere (no BVH around it), instances under static and moving transforms, an unused object
_OBJ = _WB + ('\nObjectBegin "blob"\nMaterial "plastic" "color Kd" [.2 .5 .2] "color Ks" [.4 .4 .4] "float roughness" [.05]\n'
              'Shape "sphere" "float radius" [%g]\nTranslate %g 0 0\nMaterial "matte" "color Kd" [.6 .3 .2]\n'
              'Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [%s]\nShape "disk" "float radius" [%g] "float height" [%g]\nObjectEnd\n'
              'ObjectBegin "ball"\nMaterial "mirror"\nShape "sphere" "float radius" [%g]\nObjectEnd\n'
              'ObjectBegin "unused"\nShape "sphere"\nObjectEnd\nObjectBegin "empty"\nObjectEnd\nObjectInstance "empty"\n'
              'AttributeBegin\nTranslate %s\nObjectInstance "blob"\nRotate 40 0 1 0\nTranslate %s\nScale 1.5 .7 1\nObjectInstance "blob"\nObjectInstance "ball"\nAttributeEnd\n'
              'AttributeBegin\nTranslate %s\nActiveTransform EndTime\nTranslate %s\nRotate 30 0 0 1\nActiveTransform All\nObjectInstance "ball"\nObjectInstance "blob"\nAttributeEnd\n')
INSTANCE_VARIANTS = [
    ("path-instances-bunny", "bunny", 48, 48, 4, 2, {"maxdepth": 6},
     [(_WB, _OBJ % (.04, .07, "-.05 0 -.05  .05 0 -.05  .05 .06 .05  -.05 .06 .05", .05, .03, .035, "-.1 .12 .1", ".15 0 .1", ".12 .2 -.05", "0 .08 0"))]),
    ("direct-instances-killeroo", "killeroo-simple", 48, 48, 4, 2, {"integrator": "directlighting"},
     [(_WB, _OBJ % (40, 70, "-50 0 -50  50 0 -50  50 60 50  -50 60 50", 50, 30, 35, "-100 120 100", "150 0 100", "120 200 -50", "0 80 0"))]),
]


# SURVEY 8(f) rank 4, pixel filters wider than half a pixel (every sample splats into neighbouring tasks' pixels): host front
# end + oracle, pinned on the reference binary at --ncores 1 (one worker thread runs the tasks in ascending order: the
# reference's own sums are ordered only there); the device entry points answer PBRTGPU_EUNSUPPORTED
_FILM = 'Film "image"'
FILTER_VARIANTS = [
    ("filter-gaussian", "killeroo-simple", 36, 28, 4, 1, {}, [(_FILM, 'PixelFilter "gaussian"\n' + _FILM)]),
    ("filter-mitchell", "killeroo-simple", 36, 28, 4, 1, {}, [(_FILM, 'PixelFilter "mitchell" "float B" [.2] "float C" [.4] "float xwidth" [3]\n' + _FILM)]),
    ("filter-triangle-crop", "bunny", 40, 40, 4, 1, {"crop": (0.2, 0.8, 0.1, 0.7)}, [(_FILM, 'PixelFilter "triangle" "float xwidth" [1.5] "float ywidth" [2.5]\n' + _FILM)]),
    ("filter-sinc", "killeroo-simple", 32, 32, 2, 1, {"integrator": "directlighting"}, [(_FILM, 'PixelFilter "sinc"\n' + _FILM)]),
    ("filter-widebox", "bunny", 32, 32, 4, 1, {}, [(_FILM, 'PixelFilter "box" "float xwidth" [1.5] "float ywidth" [.75]\n' + _FILM)]),
    # narrower than half a pixel: some samples touch no pixel at all (image.cpp:85-91), found by scripts/fuzz_scenes.py
    ("filter-narrowbox", "bunny", 32, 32, 4, 1, {}, [(_FILM, 'PixelFilter "box" "float xwidth" [.45] "float ywidth" [.3]\n' + _FILM)]),
]


def variant_scene(name):
    """HostScene of one of VARIANTS (thin-lens camera, maxdepth extremes, non-square crop window, shutter interval)."""
    key = ("variant", name)
    if key not in _cache:
        _, base, xres, yres, spp, ncores, kw, edits = [v for v in VARIANTS + EDGE_VARIANTS + SPECULAR_VARIANTS + METAL_VARIANTS + INFINITE_VARIANTS + ENVMAP_VARIANTS + NAMED_MATERIAL_VARIANTS + TEXTURE_VARIANTS + SPECTRUM_VARIANTS + HEIGHTFIELD_VARIANTS + DEGENERATE_UV_VARIANTS + INSTANCE_VARIANTS + FILTER_VARIANTS if v[0] == name][0]
        path = scenes.variant(base, scenes.STAGED + "/variant-%s.pbrt" % name, xres, yres, spp, **kw)
        text = open(path).read()
        for a, b in edits:
            assert a in text
            text = text.replace(a, b)
        open(path, "w").write(text)
        for fn, spec in ENV_MAPS.items():
            if fn in text:
                write_env_pfm(os.path.join(scenes.STAGED, fn), *spec)
        for fn, spec in TGA_MAPS.items():
            if fn in text:
                write_tga(os.path.join(scenes.STAGED, fn), *spec)
        _cache[key] = (pb.HostScene(path, ncores=ncores), ncores)
    return _cache[key]


def exrdiff(im1, im2):
    """The reference's own "same image" metric (tools/exrdiff.cpp:74-96) over RGB float images: channels that are zero in
    both are skipped; relative difference |a-b|/a above 0.5 % counts as small, above 5 % as big.  exrdiff without -d
    fails on any small or big difference.  Returns (smallDiff, bigDiff, avgDelta)."""
    a, b = np.asarray(im1, np.float32).ravel(), np.asarray(im2, np.float32).ravel()
    keep = ~((a == 0) & (b == 0))
    a, b = a[keep], b[keep]
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.abs(a - b) / a
    n = max(1, im1.size)
    avg1, avg2 = float(a.astype(np.float64).sum()) / n, float(b.astype(np.float64).sum()) / n
    avg_delta = (avg1 - avg2) / min(avg1, avg2) if min(avg1, avg2) != 0 else 0.0
    return int((d > .005).sum()), int((d > .05).sum()), avg_delta
