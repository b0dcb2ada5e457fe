"""Scene fixtures for the BASELINE.json configs.

The reference's example scenes (scenes/killeroo-simple.pbrt, scenes/bunny.pbrt,
geometry/*.pbrt, brdfs/mystique.brdf) are INPUT DATA, not source code.  They are
not committed: `stage()` copies them from the read-only reference checkout into
`scenes/_staged/` (git-ignored, but shipped to the GPU box by gpurun together
with the built .so files) and `variant()` derives the per-config scene files
SURVEY.md section 8(d) specifies (resolution, pixelsamples, SurfaceIntegrator
"path", .pfm output, optional crop window).
"""
import os
import re
import shutil

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(REPO, "scenes", "_staged")
REFERENCE = os.environ.get("PBRT_REFERENCE", "/root/reference")

_FILES = [
    "killeroo-simple.pbrt",
    "bunny.pbrt",
    "anim-killeroos-moving.pbrt",
    "geometry/killeroo.pbrt",
    "geometry/bunny.pbrt",
    "brdfs/mystique.brdf",
    "spds/metals/Au.eta.spd",
    "spds/metals/Au.k.spd",
    "spds/cie.stdillum.D6500.spd",
]


def stage(reference_root=REFERENCE):
    """Copy the fixture scenes out of the reference checkout (no-op if absent)."""
    src_root = os.path.join(reference_root, "scenes")
    if not os.path.isdir(src_root):
        return False
    for rel in _FILES:
        dst = os.path.join(STAGED, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst):
            shutil.copyfile(os.path.join(src_root, rel), dst)
            os.chmod(dst, 0o644)
    return True


def have_staged():
    return all(os.path.exists(os.path.join(STAGED, f)) for f in _FILES)


def _film(xres, yres, filename, crop):
    s = ('Film "image" "integer xresolution" [%d] "integer yresolution" [%d]\n'
         '    "string filename" "%s"\n' % (xres, yres, filename))
    if crop is not None:
        s += '    "float cropwindow" [%r %r %r %r]\n' % tuple(float(c) for c in crop)
    return s


def variant(base, out_path, xres, yres, spp, integrator="path", maxdepth=None,
            crop=None, filename=None, nlevels=None, strategy=None):
    """Write a config variant of staged scene `base` ('killeroo-simple' | 'bunny').

    The variant lives next to the staged base so that relative Include /
    filename paths (geometry/..., brdfs/...) keep resolving.  Returns out_path.
    """
    src = os.path.join(STAGED, base + ".pbrt")
    text = open(src).read()
    if filename is None:
        filename = os.path.splitext(os.path.basename(out_path))[0] + ".pfm"
    integ = 'SurfaceIntegrator "%s"' % integrator
    if maxdepth is not None:
        integ += ' "integer maxdepth" [%d]' % maxdepth
    if strategy is not None:
        integ += ' "string strategy" "%s"' % strategy
    sampler = 'Sampler "lowdiscrepancy" "integer pixelsamples" [%d]' % spp
    # drop the Film / Sampler / SurfaceIntegrator directives (incl. their
    # continuation lines) and re-emit ours in front of WorldBegin
    lines = text.split("\n")
    out, skipping = [], False
    for ln in lines:
        st = ln.strip()
        if re.match(r'^(Film|Sampler|SurfaceIntegrator)\b', st):
            skipping = True
            continue
        if skipping and (st.startswith('"') or st == ""):
            continue
        skipping = False
        if st.startswith("WorldBegin"):
            out.append(_film(xres, yres, filename, crop))
            out.append(sampler)
            out.append(integ)
        out.append(ln)
    text = "\n".join(out)
    if nlevels is not None:
        # "use a copy of geometry/killeroo.pbrt with "integer nlevels" [3]"
        geo_src = os.path.join(STAGED, "geometry", "killeroo.pbrt")
        geo_dst = os.path.join(STAGED, "geometry", "killeroo-nl%d.pbrt" % nlevels)
        if not os.path.exists(geo_dst):
            g = open(geo_src).read()
            g2, n = re.subn(r'"integer nlevels"\s*\[\s*\d+\s*\]', '"integer nlevels" [%d]' % nlevels, g, count=1)
            assert n == 1, "no nlevels parameter in geometry/killeroo.pbrt"
            with open(geo_dst, "w") as f:
                f.write(g2)
        text = text.replace("geometry/killeroo.pbrt", "geometry/killeroo-nl%d.pbrt" % nlevels)
    assert os.path.dirname(os.path.abspath(out_path)) == STAGED, \
        "variants must live in scenes/_staged so relative includes resolve"
    with open(out_path, "w") as f:
        f.write(text)
    return out_path


# ---- binary mesh ingestion ---------------------------------------------------------------------------
_PARAM = re.compile(r'\s*"(\w+)\s+(\w+)"\s*')


def _shape_params(text, pos):
    """the parameter list that follows a Shape directive's name at text[pos:] -> ([(type, name, raw value text)], end)"""
    params = []
    while True:
        m = _PARAM.match(text, pos)
        if not m:
            return params, pos
        p = m.end()
        if text[p] == "[":
            e = text.index("]", p) + 1
        elif text[p] == '"':
            e = text.index('"', p + 1) + 1
        else:
            e = re.compile(r"\S+").match(text, p).end()
        params.append((m.group(1), m.group(2), text[p:e]))
        pos = e


def binarize(scene_in, scene_out, min_triangles=64):
    """Rewrite a scene file so that its large trianglemesh / loopsubdiv shapes read their arrays from binary files
    (`"string binfile"`, parser.cpp LoadBinaryMesh; written next to scene_out).  Include files are expanded in place.
    Returns the number of shapes converted."""
    import numpy as np
    from . import write_binary_mesh
    base = os.path.dirname(os.path.abspath(scene_in))

    def expand(path):
        text = "\n".join(line.split("#", 1)[0] for line in open(path).read().split("\n"))
        return re.sub(r'Include\s+"([^"]+)"', lambda m: expand(os.path.join(base, m.group(1))), text)

    text = expand(scene_in)
    out, pos, n = [], 0, 0
    stem = os.path.splitext(os.path.abspath(scene_out))[0]
    for m in re.finditer(r'Shape\s+"(trianglemesh|loopsubdiv)"', text):
        params, end = _shape_params(text, m.end())
        arrays = {(t, k): v for t, k, v in params}
        idx = arrays.get(("integer", "indices"))
        P = arrays.get(("point", "P"))
        if idx is None or P is None or idx.count(" ") < 3 * min_triangles:
            continue
        num = lambda v, dt: np.array(v.strip("[]").split(), dt)
        N = arrays.get(("normal", "N"))
        uv = arrays.get(("float", "uv"), arrays.get(("float", "st")))
        binfile = "%s-mesh%d.bin" % (stem, n)
        write_binary_mesh(binfile, num(idx, np.int32), num(P, np.float32), None if N is None else num(N, np.float32),
                          None if uv is None else num(uv, np.float32))
        keep = " ".join('"%s %s" %s' % (t, k, v) for t, k, v in params
                        if (t, k) not in (("integer", "indices"), ("point", "P"), ("normal", "N"), ("float", "uv"), ("float", "st")))
        out.append(text[pos:m.end()] + ' "string binfile" "%s" %s\n' % (os.path.basename(binfile), keep))
        pos, n = end, n + 1
    out.append(text[pos:])
    with open(scene_out, "w") as f:
        f.write("".join(out))
    return n
