#!/usr/bin/env python
"""Random in-scope scenes: the unmodified reference binary vs the oracle vs the device code compiled for the host
(tests/native/device_on_host.cpp).  CPU only.  usage: fuzz_scenes.py [n] [seed0] [flags]
flags (joined with +): specular, metal (materials; metal builds the harness with -DPBRT_METAL=1), pairs (-DPBRT_PAIR_NODES=1),
crop, params ("xyz" / "spectrum" / constant-texture-valued Kd), instances (ObjectBegin / ObjectInstance), infinite (InfiniteAreaLights, constant or over a random .pfm radiance map; the harness is built with -DPBRT_INFINITE=1), filter (a random
PixelFilter on half of the scenes, harness built with -DPBRT_FILTER=1: such scenes run on one core -- the only order the
reference defines for them -- through the state machine, and EVERY pixel is compared), wave (PathIntegrator scenes run through the vertex-wavefront KERNELS -- kwave.cuh's k_wshade / k_weval / k_wtrace /
k_wcompact, launches as loops -- over a pool of 3, 17 or 4096 slots instead of the state machine)"""
import ctypes as C, os, subprocess, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import pbrt_v2_b200 as pb
from pbrt_v2_b200 import scenes
import oracle_api as oa

def fl(a): return " ".join("%.6g" % x for x in a)

def gen(rng, path, animated=None):
    L = []
    outname = os.path.splitext(os.path.basename(path))[0] + ".pfm"   # unique per scene: campaigns may run side by side
    eye = rng.uniform(-1, 1, 3); eye = eye / np.linalg.norm(eye) * rng.uniform(3, 6)
    L.append("LookAt %s  %s  0 1 0" % (fl(eye), fl(rng.uniform(-.3, .3, 3))))
    cam = 'Camera "perspective" "float fov" [%g]' % rng.uniform(30, 60)
    if rng.rand() < .3: cam += ' "float lensradius" [%g] "float focaldistance" [%g]' % (rng.uniform(.01, .1), rng.uniform(2, 6))
    L.append(cam)
    xr, yr = rng.choice([16, 24, 32]), rng.choice([16, 24])
    spp = int(rng.choice([1, 2, 4, 8]))
    film = 'Film "image" "integer xresolution" [%d] "integer yresolution" [%d] "string filename" "%s"' % (xr, yr, outname)
    if len(sys.argv) > 3 and "crop" in sys.argv[3] and rng.rand() < .5:
        a, b = sorted(rng.uniform(0, 1, 2)); c, d = sorted(rng.uniform(0, 1, 2))
        film += ' "float cropwindow" [%g %g %g %g]' % (a, max(b, a + .2), c, max(d, c + .2))
    if len(sys.argv) > 3 and "filter" in sys.argv[3] and rng.rand() < .5:
        name = rng.choice(["box", "gaussian", "mitchell", "sinc", "triangle"])
        f = 'PixelFilter "%s"' % name
        if rng.rand() < .6: f += ' "float xwidth" [%g] "float ywidth" [%g]' % (rng.uniform(.3, 3.5), rng.uniform(.3, 3.5))
        if name == "gaussian" and rng.rand() < .5: f += ' "float alpha" [%g]' % rng.uniform(.5, 4)
        if name == "mitchell" and rng.rand() < .5: f += ' "float B" [%g] "float C" [%g]' % (rng.uniform(0, 1), rng.uniform(0, 1))
        if name == "sinc" and rng.rand() < .5: f += ' "float tau" [%g]' % rng.uniform(1, 5)
        L.append(f)
    L.append(film)
    L.append('Sampler "lowdiscrepancy" "integer pixelsamples" [%d]' % spp)
    integ = rng.choice(["path", "directlighting"])
    s = 'SurfaceIntegrator "%s" "integer maxdepth" [%d]' % (integ, rng.choice([1, 2, 3, 5, 7]))
    if integ == "directlighting" and rng.rand() < .4: s += ' "string strategy" "one"'
    L.append(s)
    L.append("WorldBegin")
    if len(sys.argv) > 3 and "params" in sys.argv[3]:   # the textures kd() may name: a constant, a scale and a mix of constants
        L.append('Texture "fuzztex0" "color" "constant" "color value" [%s]' % fl(rng.uniform(0, 1, 3)))
        L.append('Texture "fuzztex1" "color" "scale" "texture tex1" "fuzztex0" "color tex2" [%s]' % fl(rng.uniform(0, 1.5, 3)))
        L.append('Texture "fuzztex2" "color" "mix" "texture tex1" "fuzztex1" "color tex2" [%s] "float amount" [%g]' % (fl(rng.uniform(0, 1, 3)), rng.uniform(0, 1)))
    if len(sys.argv) > 3 and "infinite" in sys.argv[3] and rng.rand() < .4:   # constant-radiance InfiniteAreaLight, rotated
        L.append("AttributeBegin")
        L.append("Rotate %g %s" % (rng.uniform(-180, 180), fl(rng.uniform(-1, 1, 3) + 1e-3)))
        light = 'LightSource "infinite" "color L" [%s] "integer nsamples" [%d]' % (fl(rng.uniform(0, .8, 3)), rng.choice([1, 2, 4]))
        if rng.rand() < .6:   # a radiance map: random resolution (mostly not a power of two), RGB or grey, either byte order
            from helpers import write_env_pfm, write_tga
            if rng.rand() < .7:
                mapname = os.path.splitext(os.path.basename(path))[0] + "-env.pfm"
                write_env_pfm(os.path.join(os.path.dirname(path), mapname), int(rng.randint(1, 24)), int(rng.randint(1, 14)),
                              int(rng.randint(1 << 30)), int(rng.choice([1, 3])), float(rng.choice([-1.0, 1.0, -0.5, 3.0])))
            else:   # a Targa file: grey / BGR(A) / colour-mapped, raw or RLE, any scan direction
                mapname = os.path.splitext(os.path.basename(path))[0] + "-env.tga"
                kind = str(rng.choice(["bgr", "mono", "mapped"]))
                write_tga(os.path.join(os.path.dirname(path), mapname), int(rng.randint(1, 24)), int(rng.randint(1, 14)),
                          int(rng.randint(1 << 30)), kind, bool(rng.rand() < .5), bool(rng.rand() < .5), bool(rng.rand() < .5),
                          bool(rng.rand() < .5) and kind != "mono", int(rng.randint(0, 4)))
            light += ' "string mapname" "%s"' % mapname
        L.append(light)
        L.append("AttributeEnd")
    def xform():
        t = []
        if rng.rand() < .8: t.append("Translate %s" % fl(rng.uniform(-1.2, 1.2, 3)))
        if rng.rand() < .6: t.append("Rotate %g %s" % (rng.uniform(-180, 180), fl(rng.uniform(-1, 1, 3) + 1e-3)))
        if rng.rand() < .5:
            sc = rng.uniform(.4, 1.6, 3) if rng.rand() < .5 else np.repeat(rng.uniform(.4, 1.6), 3)
            if rng.rand() < .25: sc[rng.randint(3)] *= -1     # swaps handedness
            t.append("Scale %s" % fl(sc))
        return t
    specular = len(sys.argv) > 3 and "specular" in sys.argv[3]
    metal = len(sys.argv) > 3 and "metal" in sys.argv[3]
    params = len(sys.argv) > 3 and "params" in sys.argv[3]
    def kd(hi):   # a diffuse colour in one of the ways a scene file can say it (without the flag: the draws of the earlier campaigns)
        if not params: return '"color Kd" [%s]' % fl(rng.uniform(0, hi, 3))
        k = rng.rand()
        if k < .6: return '"color Kd" [%s]' % fl(rng.uniform(0, hi, 3))
        if k < .8: return '"xyz Kd" [%s]' % fl(rng.uniform(.1, .8, 3))
        if k < .9:
            wl = rng.permutation(np.linspace(380, 780, rng.randint(2, 9)))
            return '"spectrum Kd" [%s]' % fl(np.stack([wl, rng.uniform(0, 1, len(wl))], 1).ravel())
        return '"texture Kd" "fuzztex%d"' % rng.randint(3)
    def material():
        if metal and rng.rand() < .3:
            if rng.rand() < .4: return 'Material "metal" "float roughness" [%g]' % rng.choice([.01, .05, .3])
            return 'Material "metal" "color eta" [%s] "color k" [%s] "float roughness" [%g]' % (
                fl(rng.uniform(.1, 2, 3)), fl(rng.uniform(1, 5, 3)), rng.choice([.5, .1, .02, 1e-5]))
        if specular and rng.rand() < .4:
            if rng.rand() < .5: return 'Material "mirror" "color Kr" [%s]' % fl(rng.uniform(.3, 1, 3))
            return 'Material "glass" "color Kr" [%s] "color Kt" [%s] "float index" [%g]' % (
                fl(rng.uniform(0, 1, 3)), fl(rng.uniform(0, 1, 3)), rng.choice([1.0, 1.2, 1.5, 2.4]))
        k = rng.rand()
        if k < .4: return 'Material "matte" %s' % kd(1)
        if k < .8: return 'Material "plastic" %s "color Ks" [%s] "float roughness" [%g]' % (
            kd(.6), fl(rng.uniform(0, .6, 3)), rng.choice([.5, .1, .03, 1e-5]))
        return 'Material "measured" "string filename" "brdfs/mystique.brdf"'
    def shape():
        k = rng.rand()
        if k < .35:
            r = rng.uniform(.2, .7)
            s = 'Shape "sphere" "float radius" [%g]' % r
            if rng.rand() < .4: s += ' "float zmin" [%g] "float zmax" [%g]' % (-r * rng.uniform(.2, 1), r * rng.uniform(.2, 1))
            if rng.rand() < .3: s += ' "float phimax" [%g]' % rng.uniform(90, 330)
            return s
        if k < .55:
            s = 'Shape "disk" "float radius" [%g] "float height" [%g]' % (rng.uniform(.3, .9), rng.uniform(-.3, .3))
            if rng.rand() < .4: s += ' "float innerradius" [%g]' % rng.uniform(.05, .25)
            if rng.rand() < .3: s += ' "float phimax" [%g]' % rng.uniform(90, 330)
            return s
        if len(sys.argv) > 3 and "subdiv" in sys.argv[3] and rng.rand() < .5:
            # LoopSubdiv (shapes/loopsubdiv.cpp): a closed octahedron or an open hexagonal fan (boundary rules), jittered
            if rng.rand() < .5:
                P = np.array([[1, 0, 0], [-1, 0, 0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], float) * rng.uniform(.3, .7)
                idx = [0, 2, 4, 2, 1, 4, 1, 3, 4, 3, 0, 4, 2, 0, 5, 1, 2, 5, 3, 1, 5, 0, 3, 5]
            else:
                ang = np.arange(6) * np.pi / 3
                P = np.vstack([[0, 0, 0], np.stack([np.cos(ang), np.sin(ang), 0 * ang], 1)]) * rng.uniform(.4, .8)
                idx = sum([[0, 1 + k, 1 + (k + 1) % 6] for k in range(6 if rng.rand() < .5 else 4)], [])
            P = P + rng.uniform(-.05, .05, P.shape)
            return 'Shape "loopsubdiv" "integer nlevels" [%d] "integer indices" [%s] "point P" [%s]' % (
                rng.randint(1, 4), fl(idx), fl(P.ravel()))
        nv = rng.randint(4, 9)
        P = rng.uniform(-.8, .8, (nv, 3))
        nt = rng.randint(2, 7)
        idx = np.array([rng.choice(nv, 3, replace=False) for _ in range(nt)])
        s = 'Shape "trianglemesh" "integer indices" [%s] "point P" [%s]' % (fl(idx.ravel()), fl(P.ravel()))
        if rng.rand() < .5:
            N = rng.uniform(-1, 1, (nv, 3)); N /= np.linalg.norm(N, axis=1, keepdims=True)
            s += ' "normal N" [%s]' % fl(N.ravel())
        if rng.rand() < .5: s += ' "float uv" [%s]' % fl(rng.uniform(0, 1, nv * 2))
        return s
    for _ in range(rng.randint(0, 3)):
        L.append('LightSource "point" "color I" [%s] "point from" [%s]' % (fl(rng.uniform(1, 8, 3)), fl(rng.uniform(-3, 3, 3))))
    for _ in range(rng.randint(0, 3)):
        L.append("AttributeBegin")
        L += xform()
        if rng.rand() < .3: L.append("ReverseOrientation")
        L.append('AreaLightSource "area" "color L" [%s] "integer nsamples" [%d]' % (fl(rng.uniform(2, 20, 3)), rng.choice([1, 2, 3, 4, 8])))
        L.append(material())
        L.append('Shape "sphere" "float radius" [%g]' % rng.uniform(.1, .4) if rng.rand() < .5 else
                 'Shape "disk" "float radius" [%g]' % rng.uniform(.2, .6))
        L.append("AttributeEnd")
    L.append('Material "matte" "color Kd" [.6 .6 .6]')
    L.append('Shape "trianglemesh" "integer indices" [0 1 2 2 3 0] "point P" [-4 -1.5 -4  4 -1.5 -4  4 -1.5 4  -4 -1.5 4]')
    if animated is None:
        animated = len(sys.argv) > 3 and "animated" in sys.argv[3]
    if len(sys.argv) > 3 and "instances" in sys.argv[3] and rng.rand() < .6:
        # ObjectBegin / ObjectInstance: objects of one to three shapes (one sphere or disk alone gets no aggregate), placed one
        # to three times each under static or (with "animated") moving transforms; now and then an object nobody instantiates
        names = []
        for o in range(rng.randint(1, 3)):
            L.append('ObjectBegin "obj%d"' % o)
            for _ in range(rng.randint(1, 4)):
                if rng.rand() < .5: L += xform()
                L.append(material())
                L.append(shape())
            L.append("ObjectEnd")
            names.append("obj%d" % o)
        for _ in range(rng.randint(1, 5)):
            L.append("AttributeBegin")
            L += xform()
            if animated and rng.rand() < .4:
                L.append("ActiveTransform EndTime")
                L.append("Translate %s" % fl(rng.uniform(-.5, .5, 3)))
                if rng.rand() < .5: L.append("Rotate %g %s" % (rng.uniform(-120, 120), fl(rng.uniform(-1, 1, 3) + 1e-3)))
                L.append("ActiveTransform All")
            L.append('ObjectInstance "%s"' % names[rng.randint(len(names))])
            L.append("AttributeEnd")
    for _ in range(rng.randint(2, 7)):
        L.append("AttributeBegin")
        L += xform()
        if animated and rng.rand() < .5:   # a TransformedPrimitive: the two CTMs differ (api.cpp:1011-1043)
            L.append("ActiveTransform %s" % rng.choice(["StartTime", "EndTime"]))
            k = rng.rand()
            if k < .5: L.append("Translate %s" % fl(rng.uniform(-.6, .6, 3)))
            if k > .3: L.append("Rotate %g %s" % (rng.uniform(-170, 170), fl(rng.uniform(-1, 1, 3) + 1e-3)))
            if rng.rand() < .3: L.append("Scale %s" % fl(rng.uniform(.6, 1.5, 3)))
            L.append("ActiveTransform All")
            if rng.rand() < .5: L += xform()
        if rng.rand() < .2: L.append("ReverseOrientation")
        L.append(material())
        L.append(shape())
        L.append("AttributeEnd")
    L.append("WorldEnd")
    open(path, "w").write("\n".join(L) + "\n")
    return spp

def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    scenes.stage()
    # optional third argument: kernel build flags to fuzz, e.g. "specular", "pairs", "specular+pairs"
    spec = len(sys.argv) > 3 and "specular" in sys.argv[3]
    pairs = len(sys.argv) > 3 and "pairs" in sys.argv[3]
    metal = len(sys.argv) > 3 and "metal" in sys.argv[3]
    wave = len(sys.argv) > 3 and "wave" in sys.argv[3]
    infinite = len(sys.argv) > 3 and "infinite" in sys.argv[3]
    filt = len(sys.argv) > 3 and "filter" in sys.argv[3]
    so = "/tmp/libdoh_fuzz%s%s%s%s%s.so" % ("_spec" if spec else "", "_pairs" if pairs else "", "_metal" if metal else "", "_inf" if infinite else "", "_filt" if filt else "")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden", "-Wl,-Bsymbolic",
                    "-I/usr/local/cuda/include"] + (["-DPBRT_SPECULAR=1"] if spec else []) + (["-DPBRT_PAIR_NODES=1"] if pairs else []) + (["-DPBRT_METAL=1"] if metal else []) + (["-DPBRT_INFINITE=1"] if infinite else []) + (["-DPBRT_FILTER=1"] if filt else []) + (["-DPBRT_SAFE_INVERSE=1"] if filt else []) +
                   [os.path.join(REPO, "tests/native/device_on_host.cpp"), "-o", so], check=True)
    doh = C.CDLL(so)
    doh.doh_render.argtypes = [C.c_void_p, C.POINTER(pb.RenderDesc), C.c_void_p]
    bad = 0
    for k in range(n):
        seed = seed0 + k
        rng = np.random.RandomState(seed)
        path = os.path.join(scenes.STAGED, "fuzz-%d.pbrt" % seed)
        spp = gen(rng, path)
        ncores = int(rng.choice([1, 2, 64]))
        filtered = "PixelFilter" in open(path).read()
        if filtered: ncores = 1
        quick = bool(rng.rand() < .2)   # --quick: film / 4, one pixel sample, light samples / 4
        try:
            hs = pb.HostScene(path, ncores=ncores, quick=quick)
        except pb.PbrtGpuError as e:
            print("seed %d: host front end: %s" % (seed, str(e)[:120])); continue
        try:
            ref, _ = oa.run_reference(path, ncores, timeout=120, extra_args=["--quick"] if quick else [])
        except Exception as e:
            print("seed %d: reference failed: %r" % (seed, e)); continue
        px, _ = oa.render(hs, threads=1)
        o = hs.pixels_to_rgb(px)
        rd = hs.render_desc()
        px2 = hs.new_pixels()
        doh.doh_set_vertex_form(int(rng.choice([3, 17, 4096])) if wave and not filtered else 0)
        doh.doh_render(hs.desc_ptr, C.byref(rd), px2.ctypes.data)
        own = px["weightSum"] == rd.samplesPerPixel
        if filtered: own = np.ones_like(own)
        d_ref = int((o.view(np.uint32) != ref.view(np.uint32)).any(axis=2)[own].sum()) if o.shape == ref.shape else -1
        d_dev = int((px2["Lxyz"].view(np.uint32) != px["Lxyz"].view(np.uint32)).any(axis=2)[own].sum())
        w_dev = bool(np.array_equal(px2["weightSum"].view(np.uint32), px["weightSum"].view(np.uint32)))
        flag = "" if (d_ref == 0 and d_dev == 0 and w_dev) else "   <<<<<<"
        bad += bool(flag)
        if not flag:   # keep only the scenes worth looking at
            for f in (path, os.path.splitext(path)[0] + ".pfm", os.path.splitext(path)[0] + "-env.pfm", os.path.splitext(path)[0] + "-env.tga"):
                if os.path.exists(f): os.remove(f)
        print("seed %d: %dx%d spp %d%s ncores %d integ %d  oracle-vs-reference %d  device-on-host-vs-oracle %d (weights %s)  mean %.4g%s" % (
            seed, rd.xResolution, rd.yResolution, rd.samplesPerPixel, " quick" if quick else "", ncores, rd.integrator, d_ref, d_dev, w_dev, float(ref.mean()), flag), flush=True)
    print("mismatching scenes:", bad)

if __name__ == "__main__":
    main()
