"""pbrt_v2_b200 -- Python bindings (ctypes) for the B200 path-tracing hot path.

Two in-tree shared libraries make up the product:

* ``libpbrtgpu.so``  -- hand-written sm_100a CUDA kernels behind the C ABI of
  ``include/pbrtgpu.h`` (scene upload, render, ray-batch intersect, LD samples).
* ``libpbrthost.so`` -- the C++ host side that mirrors pbrt-v2's
  Renderer / SurfaceIntegrator / Aggregate surface: scene-file front end,
  ``GpuBVHAccel`` build, flattening, ``GpuSamplerRenderer::Render``.

There is NO CPU fallback: if the libraries are missing or no CUDA device is
present the calls raise.  (The CPU restatement under ``oracle/`` is test
infrastructure and is never imported from here.)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(_HERE)


class PbrtGpuError(RuntimeError):
    pass


# ---- POD mirrors of include/pbrtgpu.h ---------------------------------------
class RenderDesc(C.Structure):
    _fields_ = [
        ("rasterToCamera", C.c_float * 16),
        ("cameraToWorld", C.c_float * 16),
        ("lensRadius", C.c_float), ("focalDistance", C.c_float),
        ("shutterOpen", C.c_float), ("shutterClose", C.c_float),
        ("xResolution", C.c_int32), ("yResolution", C.c_int32),
        ("cropWindow", C.c_float * 4),
        ("filterXWidth", C.c_float), ("filterYWidth", C.c_float),
        ("samplesPerPixel", C.c_int32),
        ("maxDepth", C.c_int32),
        ("integrator", C.c_int32), ("lightStrategy", C.c_int32),
        ("n1D", C.c_int32), ("n2D", C.c_int32),
        ("size1D", C.c_uint16 * 32), ("size2D", C.c_uint16 * 32),
        ("nTasks", C.c_int32),
        ("taskBegin", C.c_int32), ("taskEnd", C.c_int32),
        ("shardCount", C.c_int32), ("shardIndex", C.c_int32), ("shardBlock", C.c_int32),
        ("hasFilterTable", C.c_int32), ("filterTable", C.c_float * 256),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("renderMs", C.c_double), ("tableMs", C.c_double), ("traceMs", C.c_double), ("tailMs", C.c_double),
        ("cameraSamples", C.c_uint64),
        ("closestRays", C.c_uint64), ("shadowRays", C.c_uint64),
        ("nodesVisited", C.c_uint64), ("primTests", C.c_uint64),
        ("kernelLaunches", C.c_uint64),
        ("rayKernelMs", C.c_double), ("rayKernelLaunches", C.c_uint64),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SubdivMesh(C.Structure):
    """PbrtGpuSubdivMesh"""
    _fields_ = [("nVertices", C.c_uint32), ("nFaces", C.c_uint32), ("P", C.c_void_p), ("startFace", C.c_void_p),
                ("vflags", C.c_void_p), ("faceV", C.c_void_p), ("faceF", C.c_void_p)]


class BuildStats(C.Structure):
    """PbrtGpuBuildStats"""
    _fields_ = [("buildMs", C.c_double), ("nLevels", C.c_uint32), ("reserved", C.c_uint32)]


# PbrtGpuBVHNode = LinearBVHNode (accelerators/bvh.cpp:113-123)
NODE_DTYPE = np.dtype([("bmin", np.float32, 3), ("bmax", np.float32, 3), ("offset", np.uint32), ("nPrimitives", np.uint8),
                       ("axis", np.uint8), ("pad", np.uint8, 2)])
SPLIT_MIDDLE, SPLIT_EQUAL_COUNTS, SPLIT_SAH = 0, 1, 2
PIXEL_DTYPE = np.dtype([("Lxyz", np.float32, 3), ("weightSum", np.float32)])
RAY_DTYPE = np.dtype([("o", np.float32, 3), ("d", np.float32, 3), ("mint", np.float32), ("maxt", np.float32), ("time", np.float32)])
HIT_DTYPE = np.dtype([("hit", np.int32), ("primitive", np.uint32), ("t", np.float32),
                      ("b1", np.float32), ("b2", np.float32), ("p", np.float32, 3), ("n", np.float32, 3)])
SAMPLE_FLOATS = 37   # path integrator layout


def set_path_layout(rd):
    """Sample::n1D / n2D of PathIntegrator + EmissionIntegrator (path.cpp:41-49, emission.cpp:40-43)"""
    rd.integrator, rd.lightStrategy = 0, 0
    rd.n1D, rd.n2D = 14, 9
    for i in range(14):
        rd.size1D[i] = 1
    for i in range(9):
        rd.size2D[i] = 1
    return rd


def sample_floats(rd):
    """floats per camera sample of pbrtgpu_ld_samples / oracle_ld_samples for rd's Sample layout"""
    return 5 + sum(rd.size1D[i] for i in range(rd.n1D)) + 2 * sum(rd.size2D[i] for i in range(rd.n2D))

_gpu = None
_host = None


def _load(name):
    path = os.path.join(_HERE, name)
    if not os.path.exists(path):
        raise PbrtGpuError(
            "%s is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make` in the repo root); there is no CPU fallback" % path)
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


def gpu_lib():
    """libpbrtgpu.so (the CUDA hot path)."""
    global _gpu
    if _gpu is None:
        lib = _load(os.environ.get("PBRTGPU_LIB", "libpbrtgpu.so"))   # (override: kernel A/B experiments)
        lib.pbrtgpu_last_error.restype = C.c_char_p
        lib.pbrtgpu_scene_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        lib.pbrtgpu_scene_destroy.argtypes = [C.c_void_p]
        lib.pbrtgpu_scene_update.argtypes = [C.c_void_p, C.c_void_p]
        lib.pbrtgpu_scene_destroy.restype = None
        lib.pbrtgpu_render.argtypes = [C.c_void_p, C.POINTER(RenderDesc), C.c_void_p, C.POINTER(Stats)]
        lib.pbrtgpu_render_device.argtypes = [C.c_void_p, C.POINTER(RenderDesc), C.c_void_p, C.c_void_p,
                                              C.POINTER(Stats)]
        lib.pbrtgpu_intersect.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.pbrtgpu_intersect_p.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.pbrtgpu_ld_samples.argtypes = [C.POINTER(RenderDesc), C.c_int, C.c_void_p]
        lib.pbrtgpu_comm_unique_id.argtypes = [C.c_void_p]
        lib.pbrtgpu_comm_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        lib.pbrtgpu_comm_destroy.argtypes = [C.c_void_p]
        lib.pbrtgpu_comm_destroy.restype = None
        lib.pbrtgpu_render_sharded.argtypes = [C.c_void_p, C.POINTER(RenderDesc), C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        lib.pbrtgpu_set_device.argtypes = [C.c_int]
        lib.pbrtgpu_set_stats.argtypes = [C.c_void_p, C.c_int]
        lib.pbrtgpu_loopsubdiv_counts.argtypes = [C.POINTER(SubdivMesh), C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        lib.pbrtgpu_loopsubdiv_refine.argtypes = [C.POINTER(SubdivMesh), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pbrtgpu_bvh_build.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.POINTER(C.c_uint32),
                                          C.c_void_p, C.POINTER(BuildStats)]
        _gpu = lib
    return _gpu


def host_lib():
    """libpbrthost.so (scene front end + GpuBVHAccel + flattening)."""
    global _host
    if _host is None:
        gpu_lib()
        lib = _load("libpbrthost.so")
        lib.pbrthost_load.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
        lib.pbrthost_load.restype = C.c_void_p
        lib.pbrthost_load_options.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int]
        lib.pbrthost_load_options.restype = C.c_void_p
        lib.pbrthost_load_synthetic.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        lib.pbrthost_load_synthetic.restype = C.c_void_p
        lib.pbrthost_write_synthetic.argtypes = [C.c_char_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p]
        lib.pbrthost_free.argtypes = [C.c_void_p]
        lib.pbrthost_free.restype = None
        lib.pbrthost_scene_desc.argtypes = [C.c_void_p]
        lib.pbrthost_scene_desc.restype = C.c_void_p
        lib.pbrthost_render_desc.argtypes = [C.c_void_p, C.POINTER(RenderDesc)]
        lib.pbrthost_render_desc.restype = None
        lib.pbrthost_film_extent.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 4
        lib.pbrthost_film_extent.restype = None
        lib.pbrthost_pixels_to_rgb.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pbrthost_write_image.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        lib.pbrthost_render.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.POINTER(Stats)]
        lib.pbrthost_scene_bytes.argtypes = [C.c_void_p]
        lib.pbrthost_scene_bytes.restype = C.c_ulonglong
        lib.pbrthost_scene_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint)] * 4
        lib.pbrthost_scene_counts.restype = None
        lib.pbrthost_build_input.argtypes = [C.c_void_p, C.c_void_p]
        lib.pbrthost_build_input.restype = C.c_uint
        lib.pbrthost_scene_nodes.argtypes = [C.c_void_p, C.c_void_p]
        lib.pbrthost_scene_nodes.restype = None
        lib.pbrthost_loopsubdiv_topology.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 6
        lib.pbrthost_loopsubdiv_refine.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint32),
                                                   C.POINTER(C.c_uint32), C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pbrthost_bvh_build.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.POINTER(C.c_uint32),
                                           C.c_void_p]
        lib.pbrthost_envmap_build.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.pbrthost_envmap_build.restype = C.c_ulonglong
        lib.pbrthost_read_image.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
        lib.pbrthost_write_exr.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
        _host = lib
    return _host


class SceneDesc(C.Structure):
    """include/pbrtgpu.h PbrtGpuSceneDesc, for tools that inspect or rewrite a descriptor (HostScene.desc_ptr points at one)"""
    _fields_ = [("nNodes", C.c_uint32), ("nodes", C.c_void_p), ("nPrimitives", C.c_uint32), ("primitives", C.c_void_p),
                ("nTriangles", C.c_uint32), ("triangles", C.c_void_p), ("nMeshes", C.c_uint32), ("meshes", C.c_void_p),
                ("nVertices", C.c_uint32), ("P", C.c_void_p), ("N", C.c_void_p), ("uv", C.c_void_p),
                ("nQuadrics", C.c_uint32), ("quadrics", C.c_void_p), ("nXforms", C.c_uint32), ("xforms", C.c_void_p),
                ("nMaterials", C.c_uint32), ("materials", C.c_void_p), ("nMeasured", C.c_uint32), ("measured", C.c_void_p),
                ("nKdNodes", C.c_uint32), ("kdNodes", C.c_void_p), ("kdData", C.c_void_p),
                ("nLights", C.c_uint32), ("lights", C.c_void_p), ("nLightShapes", C.c_uint32), ("lightShapes", C.c_void_p),
                ("nAnimated", C.c_uint32), ("animated", C.c_void_p), ("nEnvMaps", C.c_uint32), ("envMaps", C.c_void_p),
                ("nEnvData", C.c_uint32), ("envData", C.c_void_p)]


def scene_validate(desc_ptr):
    """pbrtgpu_scene_validate: (code, message) -- the checks pbrtgpu_scene_create runs first, without a device"""
    lib = gpu_lib()
    lib.pbrtgpu_scene_validate.argtypes = [C.c_void_p]
    rc = lib.pbrtgpu_scene_validate(desc_ptr)
    return rc, (lib.pbrtgpu_last_error().decode() if rc else "")


def _check(rc, what):
    if rc != 0:
        raise PbrtGpuError("%s failed (code %d): %s" % (what, rc, gpu_lib().pbrtgpu_last_error().decode()))


class HostScene(object):
    """A parsed + flattened scene (host memory only; no GPU needed)."""

    def __init__(self, filename, ncores=0, _handle=None, quick=False):
        lib = host_lib()
        if _handle is not None:
            self._h, self.filename = _handle, filename
            return
        err = C.create_string_buffer(1024)
        self._h = lib.pbrthost_load_options(os.fsencode(filename), int(ncores), int(bool(quick)), err, len(err))
        if not self._h:
            raise PbrtGpuError("cannot load %s: %s" % (filename, err.value.decode()))
        self.filename = filename

    @classmethod
    def synthetic(cls, ntris, xres, yres, spp, maxdepth=8, ncores=0):
        """BASELINE.json configs[4] (random-triangle scene), built directly in memory."""
        err = C.create_string_buffer(1024)
        h = host_lib().pbrthost_load_synthetic(int(ntris), int(xres), int(yres), int(spp), int(maxdepth), int(ncores),
                                               err, len(err))
        if not h:
            raise PbrtGpuError("cannot build synthetic scene: %s" % err.value.decode())
        return cls("<synthetic %d triangles>" % ntris, _handle=h)

    @staticmethod
    def write_synthetic(filename, ntris, xres, yres, spp, maxdepth=8, outfile="synthetic.pfm"):
        """The same scene as pbrt text, for the unmodified reference."""
        rc = host_lib().pbrthost_write_synthetic(os.fsencode(filename), int(ntris), int(xres), int(yres), int(spp),
                                                 int(maxdepth), os.fsencode(outfile))
        if rc:
            raise PbrtGpuError("cannot write %s" % filename)
        return filename

    def close(self):
        if getattr(self, "_h", None):
            host_lib().pbrthost_free(self._h)
            self._h = None

    __del__ = close

    @property
    def desc_ptr(self):
        """const PbrtGpuSceneDesc* (valid while this object lives)."""
        return host_lib().pbrthost_scene_desc(self._h)

    def scene_bytes(self):
        return int(host_lib().pbrthost_scene_bytes(self._h))

    def counts(self):
        v = [C.c_uint() for _ in range(4)]
        host_lib().pbrthost_scene_counts(self._h, *[C.byref(x) for x in v])
        return dict(zip(("nodes", "primitives", "triangles", "lights"), (x.value for x in v)))

    def build_input(self):
        """the BVH build's input: world bounds of the refined primitives in the order BVHAccel gets them, (n, 6) float32
        pMin xyz, pMax xyz.  Only kept by scenes loaded after keep_build_input(True)."""
        n = host_lib().pbrthost_build_input(self._h, None)
        out = np.zeros((n, 6), np.float32)
        host_lib().pbrthost_build_input(self._h, out.ctypes.data)
        return out

    def nodes(self):
        """the top-level accelerator's flattened LinearBVHNode array"""
        out = np.zeros(self.counts()["nodes"], NODE_DTYPE)
        host_lib().pbrthost_scene_nodes(self._h, out.ctypes.data)
        return out

    def render_desc(self):
        rd = RenderDesc()
        host_lib().pbrthost_render_desc(self._h, C.byref(rd))
        return rd

    def film_extent(self):
        v = [C.c_int() for _ in range(4)]
        host_lib().pbrthost_film_extent(self._h, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)   # xStart, yStart, xCount, yCount

    def new_pixels(self):
        _, _, xc, yc = self.film_extent()
        return np.zeros((yc, xc), dtype=PIXEL_DTYPE)

    def pixels_to_rgb(self, pixels):
        """ImageFilm::WriteImage's conversion (film/image.cpp:178-205) -> (H, W, 3) float32."""
        _, _, xc, yc = self.film_extent()
        pixels = np.ascontiguousarray(pixels)
        rgb = np.zeros((yc, xc, 3), dtype=np.float32)
        host_lib().pbrthost_pixels_to_rgb(self._h, pixels.ctypes.data, rgb.ctypes.data)
        return rgb

    def write_image(self, pixels, filename):
        pixels = np.ascontiguousarray(pixels)
        rc = host_lib().pbrthost_write_image(self._h, pixels.ctypes.data, os.fsencode(filename))
        if rc:
            raise PbrtGpuError("cannot write %s" % filename)

    def render(self, write_image=False):
        """Renderer::Render(scene) through the C++ host class (scene upload + render + D2H)."""
        px = self.new_pixels()
        st = Stats()
        rc = host_lib().pbrthost_render(self._h, int(write_image), px.ctypes.data, C.byref(st))
        if rc:
            raise PbrtGpuError("GpuSamplerRenderer::Render failed: %s" % gpu_lib().pbrtgpu_last_error().decode())
        return px, st


class GpuScene(object):
    """Device-resident scene (pbrtgpu_scene_create)."""

    def __init__(self, host_scene):
        self.host = host_scene
        h = C.c_void_p()
        _check(gpu_lib().pbrtgpu_scene_create(host_scene.desc_ptr, C.byref(h)), "pbrtgpu_scene_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            gpu_lib().pbrtgpu_scene_destroy(self._h)
            self._h = None

    __del__ = close

    def update(self, host_scene):
        """pbrtgpu_scene_update: copy a (new) flattened scene into this handle's device buffers."""
        _check(gpu_lib().pbrtgpu_scene_update(self._h, host_scene.desc_ptr), "pbrtgpu_scene_update")
        self.host = host_scene

    def set_stats(self, on):
        """switch the ray / node / primitive-test counters on (slower kernel instantiation)."""
        _check(gpu_lib().pbrtgpu_set_stats(self._h, int(on)), "pbrtgpu_set_stats")

    def render(self, rd=None, out=None):
        """pbrtgpu_render into HOST pixels; returns (pixels, Stats)."""
        rd = rd or self.host.render_desc()
        px = out if out is not None else self.host.new_pixels()
        st = Stats()
        _check(gpu_lib().pbrtgpu_render(self._h, C.byref(rd), px.ctypes.data, C.byref(st)), "pbrtgpu_render")
        return px, st

    def render_sharded(self, rd, comm, out=None, fetch=True):
        """pbrtgpu_render_sharded: this rank's share of the tasks + the library's NCCL reduce; rank 0 gets the film
        (fetch=False: the summed film stays on the device)."""
        px = out if out is not None else (self.host.new_pixels() if (comm.rank == 0 and fetch) else None)
        st = Stats()
        _check(gpu_lib().pbrtgpu_render_sharded(self._h, C.byref(rd), comm._h, px.ctypes.data if px is not None else None,
                                                C.byref(st)), "pbrtgpu_render_sharded")
        return px, st

    def render_device(self, rd, device_ptr, stream=None, want_stats=False):
        """pbrtgpu_render_device into DEVICE memory (e.g. a torch tensor's data_ptr())."""
        st = Stats() if want_stats else None
        _check(gpu_lib().pbrtgpu_render_device(self._h, C.byref(rd), C.c_void_p(device_ptr),
                                               C.c_void_p(stream or 0),
                                               C.byref(st) if st is not None else None), "pbrtgpu_render_device")
        return st

    def intersect(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        hits = np.zeros(len(rays), dtype=HIT_DTYPE)
        _check(gpu_lib().pbrtgpu_intersect(self._h, rays.ctypes.data, len(rays), hits.ctypes.data), "pbrtgpu_intersect")
        return hits

    def intersect_p(self, rays):
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        occ = np.zeros(len(rays), dtype=np.uint8)
        _check(gpu_lib().pbrtgpu_intersect_p(self._h, rays.ctypes.data, len(rays), occ.ctypes.data), "pbrtgpu_intersect_p")
        return occ


COMM_ID_BYTES = 128


def comm_unique_id():
    """rank 0: the 128-byte id every rank needs to join the communicator (pbrtgpu_comm_unique_id)"""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    _check(gpu_lib().pbrtgpu_comm_unique_id(buf), "pbrtgpu_comm_unique_id")
    return buf.raw


class Comm(object):
    """The library's NCCL communicator (pbrtgpu_comm_create) on the current device."""

    def __init__(self, unique_id, rank, world):
        assert len(unique_id) == COMM_ID_BYTES
        h = C.c_void_p()
        _check(gpu_lib().pbrtgpu_comm_create(unique_id, int(rank), int(world), C.byref(h)), "pbrtgpu_comm_create")
        self._h, self.rank, self.world = h, int(rank), int(world)

    def close(self):
        if getattr(self, "_h", None):
            gpu_lib().pbrtgpu_comm_destroy(self._h)
            self._h = None

    __del__ = close


def ld_samples(rd, task_num):
    """LDPixelSample for the first pixel of `task_num` on the GPU -> (spp, 37) float32."""
    out = np.zeros((rd.samplesPerPixel, sample_floats(rd)), dtype=np.float32)
    _check(gpu_lib().pbrtgpu_ld_samples(C.byref(rd), int(task_num), out.ctypes.data), "pbrtgpu_ld_samples")
    return out


def _bvh_build(fn, bounds, max_prims_in_node, split_method, stats):
    bounds = np.ascontiguousarray(bounds, np.float32).reshape(-1, 6)
    n = len(bounds)
    nodes = np.zeros(max(2 * n, 1), NODE_DTYPE)
    order = np.zeros(max(n, 1), np.uint32)
    nn = C.c_uint32()
    args = [bounds.ctypes.data, n, int(max_prims_in_node), int(split_method), nodes.ctypes.data, C.byref(nn), order.ctypes.data]
    return fn(*(args + ([C.byref(stats)] if stats is not None else []))), nodes, nn, order[:n]


def bvh_build(bounds, max_prims_in_node=4, split_method=SPLIT_SAH):
    """BVHAccel's build on the GPU (pbrtgpu_bvh_build) -> (nodes, order, BuildStats)"""
    st = BuildStats()
    rc, nodes, nn, order = _bvh_build(gpu_lib().pbrtgpu_bvh_build, bounds, max_prims_in_node, split_method, st)
    _check(rc, "pbrtgpu_bvh_build")
    return nodes[:nn.value].copy(), order, st


def host_bvh_build(bounds, max_prims_in_node=4, split_method=SPLIT_SAH):
    """the same build by the host front end's recursive builder (pbrthost_bvh_build) -> (nodes, order)"""
    rc, nodes, nn, order = _bvh_build(host_lib().pbrthost_bvh_build, bounds, max_prims_in_node, split_method, None)
    if rc:
        raise PbrtGpuError("pbrthost_bvh_build: bad arguments")
    return nodes[:nn.value].copy(), order


def write_binary_mesh(path, indices, P, N=None, uv=None):
    """the `"string binfile"` form of a trianglemesh / loopsubdiv's arrays (csrc/host/parser.cpp LoadBinaryMesh)"""
    indices = np.ascontiguousarray(indices, "<i4").ravel()
    P = np.ascontiguousarray(P, "<f4").reshape(-1, 3)
    flags = (1 if N is not None else 0) | (2 if uv is not None else 0)
    with open(path, "wb") as f:
        f.write(b"PBRTMESH" + np.array([flags, len(indices), len(P)], "<u4").tobytes())
        f.write(indices.tobytes())
        f.write(P.tobytes())
        if N is not None:
            f.write(np.ascontiguousarray(N, "<f4").reshape(len(P), 3).tobytes())
        if uv is not None:
            f.write(np.ascontiguousarray(uv, "<f4").reshape(len(P), 2).tobytes())


def host_read_image(filename):
    """the host front end's ReadImage (.pfm, .exr) -> float32 [height, width, 3]"""
    w, h = C.c_int(), C.c_int()
    if host_lib().pbrthost_read_image(filename.encode(), C.byref(w), C.byref(h), None):
        raise PbrtGpuError("cannot read image %r" % filename)
    out = np.zeros((h.value, w.value, 3), np.float32)
    host_lib().pbrthost_read_image(filename.encode(), C.byref(w), C.byref(h), out.ctypes.data)
    return out


def host_write_exr(filename, rgb):
    """the host front end's WriteImageEXR on a float32 [height, width, 3] image"""
    rgb = np.ascontiguousarray(rgb, np.float32)
    if host_lib().pbrthost_write_exr(filename.encode(), rgb.ctypes.data, rgb.shape[1], rgb.shape[0]):
        raise PbrtGpuError("cannot write %r" % filename)


def host_envmap_build(pfm, L):
    """InfiniteAreaLight's constructor by the host front end (csrc/host/envmap.cpp) -> dict of the PbrtGpuEnvMap arrays:
    width, height, nu, nv, texels [height, width, 3], func [nv, nu], cdf [nv, nu + 1], funcInt [nv], margCdf [nv + 1],
    margFuncInt"""
    L = np.ascontiguousarray(L, np.float32)
    head = np.zeros(4, np.uint32)
    name = pfm.encode() if pfm else None
    n = host_lib().pbrthost_envmap_build(name, L.ctypes.data, head.ctypes.data, None)
    if not n:
        raise PbrtGpuError("pbrthost_envmap_build: cannot read %r" % pfm)
    buf = np.zeros(n, np.float32)
    host_lib().pbrthost_envmap_build(name, L.ctypes.data, head.ctypes.data, buf.ctypes.data)
    return parse_envmap_dump(head, buf)


def parse_envmap_dump(head, buf):
    """the layout shared by pbrthost_envmap_build and `refdump envmap`"""
    w, h, nu, nv = [int(x) for x in head]
    out, at = {"width": w, "height": h, "nu": nu, "nv": nv}, 0
    for key, shape in (("texels", (h, w, 3)), ("func", (nv, nu)), ("cdf", (nv, nu + 1)), ("funcInt", (nv,)), ("margCdf", (nv + 1,)),
                       ("margFuncInt", (1,))):
        n = int(np.prod(shape))
        out[key] = buf[at:at + n].reshape(shape).copy()
        at += n
    assert at == len(buf)
    return out


class ControlMesh(object):
    """A LoopSubdiv control mesh with the constructor's topology (loopsubdiv.cpp:156-216) as PbrtGpuSubdivMesh arrays."""

    def __init__(self, indices, P):
        self.indices = np.ascontiguousarray(indices, np.int32).reshape(-1, 3)
        self.P = np.ascontiguousarray(P, np.float32).reshape(-1, 3)
        nv, nf = len(self.P), len(self.indices)
        self.startFace = np.zeros(nv, np.int32)
        self.vflags = np.zeros(nv, np.uint8)
        self.faceV = np.zeros((nf, 3), np.int32)
        self.faceF = np.zeros((nf, 3), np.int32)
        host_lib().pbrthost_loopsubdiv_topology(nv, nf, self.indices.ctypes.data, self.P.ctypes.data, self.startFace.ctypes.data,
                                                self.vflags.ctypes.data, self.faceV.ctypes.data, self.faceF.ctypes.data)
        self.desc = SubdivMesh(nv, nf, self.P.ctypes.data, self.startFace.ctypes.data, self.vflags.ctypes.data,
                               self.faceV.ctypes.data, self.faceF.ctypes.data)

    def counts(self, nlevels):
        """(vertices, triangles) of the refined mesh (pbrtgpu_loopsubdiv_counts: validation + arithmetic, no device needed)"""
        nv, nt = C.c_uint32(), C.c_uint32()
        _check(gpu_lib().pbrtgpu_loopsubdiv_counts(C.byref(self.desc), int(nlevels), C.byref(nv), C.byref(nt)), "pbrtgpu_loopsubdiv_counts")
        return nv.value, nt.value

    def refine(self, nlevels):
        """LoopSubdiv::Refine on the GPU -> (P limit, N limit, vertex indices)"""
        nv, nt = self.counts(nlevels)
        P, N, idx = np.zeros((nv, 3), np.float32), np.zeros((nv, 3), np.float32), np.zeros((nt, 3), np.int32)
        _check(gpu_lib().pbrtgpu_loopsubdiv_refine(C.byref(self.desc), int(nlevels), P.ctypes.data, N.ctypes.data, idx.ctypes.data),
               "pbrtgpu_loopsubdiv_refine")
        return P, N, idx

    def host_refine(self, nlevels):
        """the same by the host front end's LoopSubdiv::Refine"""
        nv, nt = C.c_uint32(), C.c_uint32()
        args = [len(self.P), len(self.indices), self.indices.ctypes.data, self.P.ctypes.data, int(nlevels), C.byref(nv), C.byref(nt)]
        if host_lib().pbrthost_loopsubdiv_refine(*(args + [None, None, None])):
            raise PbrtGpuError("pbrthost_loopsubdiv_refine failed")
        P, N, idx = np.zeros((nv.value, 3), np.float32), np.zeros((nv.value, 3), np.float32), np.zeros((nt.value, 3), np.int32)
        host_lib().pbrthost_loopsubdiv_refine(*(args + [P.ctypes.data, N.ctypes.data, idx.ctypes.data]))
        return P, N, idx


def keep_build_input(on=True):
    host_lib().pbrthost_keep_build_input(int(on))


def set_bvh_parallel_min(n=1 << 17):
    """host BVH builds of at least n primitives fork their subtrees onto several threads (same tree either way)"""
    host_lib().pbrthost_set_bvh_parallel_min(int(n))


def device_count():
    return gpu_lib().pbrtgpu_device_count()
