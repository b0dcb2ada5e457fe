"""ctypes access to the oracle (oracle/liboracle.so) and the reference binary
(oracle/_ref/bin/pbrt).  TEST-ONLY: nothing under pbrt_v2_b200/ imports this."""
import ctypes as C
import os
import subprocess

import numpy as np

import pbrt_v2_b200 as pb

REPO = pb.REPO
ORACLE_SO = os.path.join(REPO, "oracle", "liboracle.so")
REF_BIN = os.path.join(REPO, "oracle", "_ref", "bin", "pbrt")


class OracleCounters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("cameraSamples", "closestRays", "shadowRays", "nodesVisited",
                                          "primTests", "shadedHitsWithNormals", "rngDraws")]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(ORACLE_SO)
        l.oracle_render.argtypes = [C.c_void_p, C.POINTER(pb.RenderDesc), C.c_void_p, C.c_int, C.POINTER(OracleCounters)]
        l.oracle_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        l.oracle_intersect_p.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        l.oracle_ld_samples.argtypes = [C.POINTER(pb.RenderDesc), C.c_int, C.c_void_p]
        l.oracle_aggregate_test.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        l.oracle_task_window.argtypes = [C.POINTER(pb.RenderDesc), C.c_int, C.POINTER(C.c_int * 4)]
        _lib = l
    return _lib


def have_oracle():
    return os.path.exists(ORACLE_SO)


def have_reference():
    return os.path.exists(REF_BIN)


def render(host_scene, rd=None, threads=None):
    rd = rd or host_scene.render_desc()
    px = host_scene.new_pixels()
    cnt = OracleCounters()
    threads = threads or os.cpu_count() or 1
    rc = lib().oracle_render(host_scene.desc_ptr, C.byref(rd), px.ctypes.data, threads, C.byref(cnt))
    assert rc == 0
    return px, cnt


def intersect(host_scene, rays):
    rays = np.ascontiguousarray(rays, dtype=pb.RAY_DTYPE)
    hits = np.zeros(len(rays), dtype=pb.HIT_DTYPE)
    lib().oracle_intersect(host_scene.desc_ptr, rays.ctypes.data, len(rays), hits.ctypes.data)
    return hits


def intersect_p(host_scene, rays):
    rays = np.ascontiguousarray(rays, dtype=pb.RAY_DTYPE)
    occ = np.zeros(len(rays), dtype=np.uint8)
    lib().oracle_intersect_p(host_scene.desc_ptr, rays.ctypes.data, len(rays), occ.ctypes.data)
    return occ


def aggregate_test(host_scene, n):
    """renderers/aggregatetest.cpp restated: returns (disagreements BVH vs exhaustive, rays, hit, t)"""
    rays = np.zeros(n, dtype=pb.RAY_DTYPE)
    hit = np.zeros(n, dtype=np.uint8)
    t = np.zeros(n, dtype=np.float32)
    bad = lib().oracle_aggregate_test(host_scene.desc_ptr, int(n), rays.ctypes.data, hit.ctypes.data, t.ctypes.data)
    return bad, rays, hit, t


def bxdf_test(exponent, wo, wi, u):
    """oracle BxDF restatements on given direction pairs -> (n, 15) float32 (see oracle_bxdf_test)"""
    wo, wi, u = [np.ascontiguousarray(a, np.float32) for a in (wo, wi, u)]
    out = np.zeros((len(wo), 15), np.float32)
    l = lib()
    l.oracle_bxdf_test.argtypes = [C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    l.oracle_bxdf_test(float(exponent), len(wo), wo.ctypes.data, wi.ctypes.data, u.ctypes.data, out.ctypes.data)
    return out


def ld_samples(rd, task_num, pixel_index=0):
    """LDPixelSample's values for pixel `pixel_index` of the task (the pixels before it sampled first, no integrator draws)"""
    out = np.zeros((rd.samplesPerPixel, pb.sample_floats(rd)), dtype=np.float32)
    f = lib().oracle_ld_samples_seq
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    rc = f(C.addressof(rd), int(task_num), int(pixel_index), out.ctypes.data)
    assert rc == 0, "empty task window or pixel outside it"
    return out


def task_window(rd, task_num):
    w = (C.c_int * 4)()
    lib().oracle_task_window(C.byref(rd), int(task_num), C.byref(w))
    return tuple(w)


def read_pfm(path):
    with open(path, "rb") as f:
        assert f.readline().strip() == b"PF"
        w, h = [int(x) for x in f.readline().split()]
        scale = float(f.readline())
        data = np.frombuffer(f.read(), dtype="<f4" if scale < 0 else ">f4", count=3 * w * h)
    return data.reshape(h, w, 3)[::-1].copy()   # rows are stored bottom-to-top


def run_reference(scene_path, ncores, timeout=3600, extra_args=()):
    """Run the UNMODIFIED reference on `scene_path` (cwd = scene dir); returns (rgb, seconds)."""
    import re, time
    cwd = os.path.dirname(os.path.abspath(scene_path))
    text = open(scene_path).read()
    out = re.search(r'"string filename"\s*"([^"]+)"', text).group(1)
    outp = os.path.join(cwd, out)
    if os.path.exists(outp):
        os.remove(outp)
    t0 = time.time()
    subprocess.run([REF_BIN, "--ncores", str(ncores), "--quiet"] + list(extra_args) + [os.path.basename(scene_path)], cwd=cwd,
                   check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout)
    dt = time.time() - t0
    return read_pfm(outp), dt
