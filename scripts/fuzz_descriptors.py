#!/usr/bin/env python
"""Malformed scene descriptors at the C ABI: "a descriptor that passes pbrtgpu_scene_validate keeps every index the kernels
follow inside its array, every walk finite and inside its stack" (include/pbrtgpu.h), tested.  Valid descriptors (the host
front end's, of small random scenes) are copied array by array, a few 32-bit words are overwritten (small and huge integers,
off-by-ones of the array sizes, NaN / infinity / denormal bit patterns), and
  * pbrtgpu_scene_validate must answer without crashing;
  * whenever it answers PBRTGPU_OK, the device code compiled for the host (tests/native/device_on_host.cpp, built here with
    AddressSanitizer: every array is a malloc'd block with red zones) renders a small film over the mutated descriptor --
    state machine, vertex form and wave kernels -- under a watchdog: an out-of-bounds read or a walk that does not end is
    a hole in the validation.
CPU only.  usage: fuzz_descriptors.py [n] [seed] [rd]     (re-executes itself with libasan preloaded; "rd": mutate
the render descriptor only)"""
import ctypes as C, os, subprocess, sys, tempfile, signal
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "scripts"))

if "libasan" not in os.environ.get("LD_PRELOAD", ""):
    asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1:allocator_may_return_null=1")
    os.execve(sys.executable, [sys.executable] + sys.argv, env)

import numpy as np
import pbrt_v2_b200 as pb
import fuzz_scenes

ELEMS = ["PbrtGpuBVHNode", "PbrtGpuPrimitive", "PbrtGpuTriangle", "PbrtGpuMesh", "PbrtGpuQuadric", "PbrtGpuXform", "PbrtGpuMaterial",
         "PbrtGpuMeasured", "PbrtGpuKdNode", "PbrtGpuBrdfSample", "PbrtGpuLight", "PbrtGpuLightShape", "PbrtGpuAnimated", "PbrtGpuEnvMap"]


SceneDesc = pb.SceneDesc


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    only_rd = len(sys.argv) > 3 and sys.argv[3] == "rd"      # leave the scene alone, overwrite words of the render descriptor only
    tmp = tempfile.mkdtemp(prefix="pbrt_fuzz_desc")
    # element sizes and sizeof(PbrtGpuSceneDesc) from the header itself
    src = os.path.join(tmp, "sizes.c")
    open(src, "w").write('#include <stdio.h>\n#include "pbrtgpu.h"\nint main(void){printf("%zu", sizeof(PbrtGpuSceneDesc));' +
                         "".join('printf(" %%zu", sizeof(%s));' % e for e in ELEMS) + "return 0;}\n")
    subprocess.run(["gcc", "-I" + os.path.join(REPO, "include"), src, "-o", os.path.join(tmp, "sizes")], check=True)
    sizes = [int(x) for x in subprocess.run([os.path.join(tmp, "sizes")], capture_output=True, text=True).stdout.split()]
    assert sizes[0] == C.sizeof(SceneDesc), (sizes[0], C.sizeof(SceneDesc))
    sz = dict(zip(ELEMS, sizes[1:]))
    arrays = [("nodes", "nNodes", sz["PbrtGpuBVHNode"]), ("primitives", "nPrimitives", sz["PbrtGpuPrimitive"]),
              ("triangles", "nTriangles", sz["PbrtGpuTriangle"]), ("meshes", "nMeshes", sz["PbrtGpuMesh"]),
              ("P", "nVertices", 12), ("N", "nVertices", 12), ("uv", "nVertices", 8),
              ("quadrics", "nQuadrics", sz["PbrtGpuQuadric"]), ("xforms", "nXforms", sz["PbrtGpuXform"]),
              ("materials", "nMaterials", sz["PbrtGpuMaterial"]), ("measured", "nMeasured", sz["PbrtGpuMeasured"]),
              ("kdNodes", "nKdNodes", sz["PbrtGpuKdNode"]), ("kdData", "nKdNodes", sz["PbrtGpuBrdfSample"]),
              ("lights", "nLights", sz["PbrtGpuLight"]), ("lightShapes", "nLightShapes", sz["PbrtGpuLightShape"]),
              ("animated", "nAnimated", sz["PbrtGpuAnimated"]), ("envMaps", "nEnvMaps", sz["PbrtGpuEnvMap"]), ("envData", "nEnvData", 4)]
    # the device code for the host, with AddressSanitizer; every gated feature on (what the f4 library holds)
    so = os.path.join(tmp, "libdoh_asan.so")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-fvisibility=hidden", "-Wl,-Bsymbolic",
                    "-fsanitize=address", "-I/usr/local/cuda/include", "-DPBRT_METAL=1", "-DPBRT_INFINITE=1", "-DPBRT_FILTER=1",
                    "-DPBRT_SAFE_INVERSE=%d" % (0 if os.environ.get("FUZZ_SHIPPED_INVERSE") else 1),   # (the shipped kernels: 0, see danim.cuh)
                    os.path.join(REPO, "tests/native/device_on_host.cpp"), "-o", so], check=True)
    doh = C.CDLL(so)
    doh.doh_render.argtypes = [C.c_void_p, C.POINTER(pb.RenderDesc), C.c_void_p]
    # the validator of an f4 build (PBRT_INFINITE etc. change what it accepts): built on demand, else the shipped library's
    f4 = os.path.join(REPO, "pbrt_v2_b200", "ab", "libpbrtgpu_f4.so")
    lib = C.CDLL(f4) if os.path.exists(f4) else pb.gpu_lib()
    print("validator:", "f4 build" if os.path.exists(f4) else "shipped build (scenes with metal / infinite lights are refused, not walked)")
    lib.pbrtgpu_scene_validate.argtypes = [C.c_void_p]
    lib.pbrtgpu_render_validate_scene.argtypes = [C.c_void_p, C.POINTER(pb.RenderDesc)]
    lib.pbrtgpu_film_extent.argtypes = [C.POINTER(pb.RenderDesc), C.c_void_p]

    sys.argv = sys.argv[:1] + ["0", "0", "specular+crop+animated+instances" if only_rd else "specular+metal+crop+infinite+animated+instances+params"]
    rs = np.random.RandomState(seed)
    from pbrt_v2_b200 import scenes
    scenes.stage()

    def watchdog(signum, frame):
        print("WATCHDOG: a walk over a validated descriptor did not end (mutation kept in", tmp, ")"); os._exit(3)
    signal.signal(signal.SIGALRM, watchdog)
    stats = {"refused": 0, "rendered": 0}
    k = 0
    while k < n:
        path = os.path.join(scenes.STAGED, "fuzzdesc-%d.pbrt" % (seed * 100000 + k))
        fuzz_scenes.gen(np.random.RandomState(700000 + seed * 100000 + k), path)
        try:
            hs = pb.HostScene(path, ncores=1)
        except pb.PbrtGpuError:
            os.remove(path); k += 1; continue
        src_desc = SceneDesc.from_address(hs.desc_ptr)
        rd = hs.render_desc()
        for _ in range(25):     # mutations of this scene
            bufs, d = {}, SceneDesc()
            C.memmove(C.byref(d), hs.desc_ptr, C.sizeof(SceneDesc))
            for field, count, esz in arrays:
                ptr, cnt = getattr(src_desc, field), getattr(src_desc, count)
                if not ptr or not cnt:
                    continue
                a = np.empty(cnt * esz // 4, np.uint32)      # (exactly the array: the red zone starts right behind it)
                C.memmove(a.ctypes.data, ptr, cnt * esz)
                bufs[field] = a
                setattr(d, field, a.ctypes.data)
            for _m in range(0 if only_rd else rs.randint(1, 4)):
                field = list(bufs)[rs.randint(len(bufs))]
                a = bufs[field]
                i = rs.randint(len(a))
                cnts = [getattr(src_desc, c) for _, c, _ in arrays]
                choice = rs.randint(6)
                if choice == 0: a[i] = rs.randint(0, 8)
                elif choice == 1: a[i] = rs.choice(cnts) + rs.randint(-1, 2) & 0xffffffff
                elif choice == 2: a[i] = rs.choice([0xffffffff, 0x7fffffff, 0x80000000, 0xfffffffe, 1 << 16, 1 << 24])
                elif choice == 3: a[i] = rs.choice([0x7fc00000, 0x7f800000, 0xff800000, 0x00000001, 0x80000000, 0x7f7fffff])   # NaN, +-inf, denormal, -0, max
                elif choice == 4: a[i] = a[rs.randint(len(a))]
                else: a[i] ^= 1 << rs.randint(32)
            if not only_rd and rs.rand() < .15:      # ... or one of the counts themselves, downwards (upwards would lie about the allocation)
                _, count, _ = arrays[rs.randint(len(arrays))]
                if getattr(d, count) > 1: setattr(d, count, rs.randint(1, getattr(d, count)))
            rc = lib.pbrtgpu_scene_validate(C.addressof(d))
            if rc != 0:
                stats["refused"] += 1
                continue
            # ... and, on a third of the runs, the render descriptor: any of its words (the head: camera, film, sampler,
            # integrator, task range; the sample-array sizes; the shard fields; the filter table)
            rdm = pb.RenderDesc()
            C.memmove(C.byref(rdm), C.byref(rd), C.sizeof(pb.RenderDesc))
            if only_rd or rs.rand() < .33:
                w = np.frombuffer((C.c_uint32 * (C.sizeof(pb.RenderDesc) // 4)).from_address(C.addressof(rdm)), np.uint32)
                for _m in range(rs.randint(1, 3)):
                    i = rs.randint(len(w)) if rs.rand() < .3 else rs.randint(32, 90)   # (the words after the two matrices are the interesting ones)
                    choice = rs.randint(5)
                    if choice == 0: w[i] = rs.randint(0, 70)
                    elif choice == 1: w[i] = rs.choice([0xffffffff, 0x7fffffff, 0x80000000, 1 << 16, 1 << 13, 8192, 4096])
                    elif choice == 2: w[i] = rs.choice([0x7fc00000, 0x7f800000, 0xff800000, 0x3f800001, 0xbf800000, 0x3f000000, 0x3e800000])
                    elif choice == 3: w[i] = w[rs.randint(32, 90)]
                    else: w[i] ^= 1 << rs.randint(32)
                if lib.pbrtgpu_render_validate_scene(C.addressof(d), C.byref(rdm)) != 0:
                    stats["refused"] += 1
                    continue
            ext = (C.c_int32 * 4)()
            assert lib.pbrtgpu_film_extent(C.byref(rdm), ext) == 0
            if ext[2] * ext[3] * rdm.samplesPerPixel > 300000 or ext[2] * ext[3] > 1 << 20:
                stats["skipped (valid, too costly here)"] = stats.get("skipped (valid, too costly here)", 0) + 1
                continue
            px = np.zeros(ext[2] * ext[3], pb.PIXEL_DTYPE)     # what a caller sizes `pixels` with
            for form in (0, 1, 64):
                doh.doh_set_vertex_form(form)
                signal.alarm(180)
                doh.doh_render(C.addressof(d), C.byref(rdm), px.ctypes.data)
                signal.alarm(0)
            doh.doh_set_vertex_form(0)
            stats["rendered"] += 1
        os.remove(path)
        k += 1
        if k % 50 == 0:
            print(k, stats, flush=True)
    print("scenes:", n, stats, "-- no sanitizer report, no runaway walk")


if __name__ == "__main__":
    main()
