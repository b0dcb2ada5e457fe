"""Host-side logic of the drop-in (no GPU): task decomposition, scene front end
errors, flattened-scene invariants, multi-rank sharding (gloo, world_size 2)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import pbrt_v2_b200 as pb
from pbrt_v2_b200 import scenes
import oracle_api as oa
from helpers import host_scene


def _round_up_pow2(v):
    return 1 << (v - 1).bit_length()


@pytest.mark.parametrize("res,ncores", [(32, 1), (32, 32), (400, 8), (100, 3)])
def test_task_count_follows_reference_rule(staged, res, ncores):
    # samplerrenderer.cpp:203-205: nTasks = RoundUpPow2(max(32*cores, nPixels/256))
    hs = host_scene("bunny", res, res, 4, ncores)
    rd = hs.render_desc()
    assert rd.nTasks == _round_up_pow2(max(32 * ncores, res * res // 256))
    assert (rd.taskBegin, rd.taskEnd) == (0, rd.nTasks)
    assert rd.maxDepth == 5 and rd.samplesPerPixel == 4


def test_task_windows_tile_the_image(staged):
    hs = host_scene("bunny", 100, 100, 4, 3)      # 128 tasks over a 100x100 image: ragged tiles
    rd = hs.render_desc()
    cover = np.zeros((100, 100), np.int32)
    for t in range(rd.nTasks):
        x0, x1, y0, y1 = oa.task_window(rd, t)
        cover[y0:y1, x0:x1] += 1
    assert (cover == 1).all()


def test_pixelsamples_round_up_to_power_of_two(staged):
    p = scenes.variant("bunny", os.path.join(scenes.STAGED, "t-spp5.pbrt"), 16, 16, 5)
    assert pb.HostScene(p, ncores=1).render_desc().samplesPerPixel == 8   # LDSampler ctor, lowdiscrepancy.cpp:40-51


def test_crop_window_extent(staged):
    hs = host_scene("bunny", 64, 64, 4, 8, crop=(0.25, 0.75, 0.5, 1.0), tag="-crop")
    assert hs.film_extent() == (16, 32, 32, 32)   # image.cpp:48-51


@pytest.mark.parametrize("text,needle", [
    ('SurfaceIntegrator "whitted"', "outside the scope"),
    ('Sampler "random"', "out of scope"),
    ('Material "uber"', "outside the hot-path scope"),
    ('Material "metal" "frobnicate eta" [5500 1]', "not supported by this front end"),
    ('Texture "t" "color" "imagemap" "string filename" "x.exr"', "outside the hot-path scope"),   # (constant / scale / mix fold into the material)
    ('Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0] "float alpha" [0]', "alpha textures"),
    ('Material "matte" "texture bumpmap" "b"', "bump maps"),
    ('Shape "trianglemesh" "integer indices" [0 1 2] "point P" [0 0 0 1 0 0 0 1 0] "vector S" [1 0 0 1 0 0 1 0 0]', "tangents"),
])
def test_out_of_scope_directives_fail_loudly(staged, text, needle):
    src = open(scenes.variant("bunny", os.path.join(scenes.STAGED, "t-err.pbrt"), 8, 8, 1)).read()
    if text.startswith("SurfaceIntegrator"):
        src = src.replace('SurfaceIntegrator "path"', text)
    elif text.startswith("Sampler"):
        src = src.replace('Sampler "lowdiscrepancy"', text)
    else:
        src = src.replace("WorldBegin", "WorldBegin\n" + text + "\n")
    p = os.path.join(scenes.STAGED, "t-err2.pbrt")
    open(p, "w").write(src)
    with pytest.raises(pb.PbrtGpuError) as e:
        pb.HostScene(p, ncores=1)
    assert needle in str(e.value)


def test_scene_preparation_options_are_parsed(staged, capfd):
    """Accelerator "bvh" "string builder" and Shape "loopsubdiv" "string refiner": "host" is the default and gives the
    same scene; an unknown value warns and uses the host; "gpu" without a device fails the load (no silent fallback)"""
    src = open(scenes.variant("killeroo-simple", os.path.join(scenes.STAGED, "t-opt.pbrt"), 8, 8, 1)).read()
    assert 'Accelerator' not in src   # (the shipped scenes leave the accelerator at its default)

    def load(accel, env=None):
        p = os.path.join(scenes.STAGED, "t-opt2.pbrt")
        open(p, "w").write(src.replace("WorldBegin", accel + "\nWorldBegin", 1))
        return pb.HostScene(p, ncores=1)
    base = load("")
    for accel in ('Accelerator "bvh" "string builder" "host"', 'Accelerator "bvh" "string builder" "fpga"'):
        hs = load(accel)
        assert hs.nodes().tobytes() == base.nodes().tobytes()
    assert 'BVH builder "fpga" unknown' in capfd.readouterr().err
    if pb.device_count() == 0:
        with pytest.raises(pb.PbrtGpuError, match="no CUDA device"):
            load('Accelerator "bvh" "string builder" "gpu"')
    # middle / equal splits are host-only: asking for the device build there warns and builds on the host
    hs = load('Accelerator "bvh" "string splitmethod" "middle" "string builder" "gpu"')
    assert hs.counts()["nodes"] > 0
    assert "only splitmethod" in capfd.readouterr().err


def test_flattened_bvh_invariants(staged):
    """depth-first layout: left child = node+1, second child > node; every primitive in exactly one leaf."""
    hs = host_scene("killeroo-simple", 8, 8, 1, 1)
    n = hs.counts()
    nodes_ptr = C.cast(hs.desc_ptr + 8, C.POINTER(C.c_void_p))[0]
    raw = np.ctypeslib.as_array(C.cast(nodes_ptr, C.POINTER(C.c_uint8)), shape=(n["nodes"] * 32,)).copy()
    node = raw.view(np.dtype([("bmin", "<f4", 3), ("bmax", "<f4", 3), ("offset", "<u4"), ("nprims", "u1"),
                              ("axis", "u1"), ("pad", "u1", 2)]))
    leaf = node["nprims"] > 0
    assert leaf.sum() + (~leaf).sum() == n["nodes"] and (~leaf).sum() == leaf.sum() - 1
    assert node["nprims"][leaf].sum() == n["primitives"]
    idx = np.arange(n["nodes"])
    assert (node["offset"][~leaf] > idx[~leaf] + 1).all()
    assert (node["axis"][~leaf] <= 2).all() and node["nprims"].max() <= 4   # maxnodeprims default, bvh.cpp:509
    first = np.sort(node["offset"][leaf])
    assert first[0] == 0 and (np.diff(first) > 0).all()
    assert (node["bmin"] <= node["bmax"]).all()


def test_sharded_oracle_render_sums_to_full(staged):
    hs = host_scene("bunny", 24, 24, 4, 18)   # 1024 tasks > 576 pixels
    full, _ = oa.render(hs, threads=2)
    from pbrt_v2_b200.multigpu import shard_render_desc
    acc = None
    for r in range(3):
        rd = shard_render_desc(hs.render_desc(), r, 3, 24, rows_per_block=2)
        part, _ = oa.render(hs, rd, threads=2)
        acc = part if acc is None else np.rec.fromarrays([acc["Lxyz"] + part["Lxyz"], acc["weightSum"] + part["weightSum"]],
                                                         dtype=pb.PIXEL_DTYPE)
    assert np.array_equal(acc["weightSum"], full["weightSum"])
    np.testing.assert_allclose(acc["Lxyz"], full["Lxyz"], rtol=1e-6)


def test_sharded_render_with_a_wide_filter_sums_to_full(staged):
    """every sample of a wide filter splats into other tasks' -- other shards' -- pixels: the shards' films still add up to the
    one-process film, to the rounding of the sums (the reduce is a plain sum of films; ncclReduce in the library)"""
    from helpers import variant_scene
    from pbrt_v2_b200.multigpu import shard_render_desc
    hs, _ = variant_scene("filter-mitchell")
    full, _ = oa.render(hs, threads=1)
    acc = None
    for r in range(4):
        rd = shard_render_desc(hs.render_desc(), r, 4, hs.render_desc().xResolution, rows_per_block=2)
        part, _ = oa.render(hs, rd, threads=1)
        acc = part if acc is None else np.rec.fromarrays([acc["Lxyz"] + part["Lxyz"], acc["weightSum"] + part["weightSum"]],
                                                         dtype=pb.PIXEL_DTYPE)
    for f in ("weightSum", "Lxyz"):
        assert np.abs(acc[f].astype(np.float64) - full[f]).max() <= 4e-6 * np.abs(full[f]).max(), f


_GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import pbrt_v2_b200 as pb, oracle_api as oa
from pbrt_v2_b200.multigpu import shard_render_desc, reduce_film
from helpers import host_scene
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
hs = host_scene("bunny", 24, 24, 4, 18)
rd = shard_render_desc(hs.render_desc(), rank, world, 24, rows_per_block=2)
px, _ = oa.render(hs, rd, threads=1)            # the CPU oracle stands in for the GPU render in this host-logic test
film = torch.from_numpy(np.concatenate([px["Lxyz"], px["weightSum"][..., None]], axis=2).copy())
reduce_film(film, world)
if rank == 0:
    full, _ = oa.render(hs, threads=1)
    want = np.concatenate([full["Lxyz"], full["weightSum"][..., None]], axis=2)
    assert np.array_equal(film.numpy()[..., 3], want[..., 3])
    np.testing.assert_allclose(film.numpy(), want, rtol=1e-6)
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_world_size_2_gloo_shard_and_reduce(staged, tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(_GLOO_WORKER)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(w), pb.REPO],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert "GLOO_OK" in out.stdout, out.stdout[-2000:]


@pytest.mark.parametrize("base", ["bunny", "killeroo-simple"])
def test_binary_mesh_ingestion_builds_the_same_scene(staged, base):
    """`"string binfile"` (mesh arrays from a binary file instead of the parameter list; trianglemesh for
    the bunny, loopsubdiv control meshes for the killeroo): same flattened scene -- node array bit for bit -- and the same film"""
    src = scenes.variant(base, os.path.join(scenes.STAGED, "binmesh-src-%s.pbrt" % base), 20, 20, 2)
    dst = os.path.join(scenes.STAGED, "binmesh-%s.pbrt" % base)
    assert scenes.binarize(src, dst) >= 1
    assert '"string binfile"' in open(dst).read()
    a, b = pb.HostScene(src, ncores=2), pb.HostScene(dst, ncores=2)
    assert a.counts() == b.counts() and a.scene_bytes() == b.scene_bytes()
    assert np.array_equal(a.nodes().view(np.uint8), b.nodes().view(np.uint8))
    pa, _ = oa.render(a, threads=4)
    pbin, _ = oa.render(b, threads=4)
    assert np.array_equal(pa.view(np.uint8), pbin.view(np.uint8))


def test_malformed_binary_mesh_is_refused(staged, tmp_path):
    scene = os.path.join(scenes.STAGED, "binmesh-bad.pbrt")
    text = open(scenes.variant("bunny", scene, 8, 8, 1)).read()
    text = text.replace("WorldEnd", 'Shape "trianglemesh" "string binfile" "%s"\nWorldEnd' % str(tmp_path / "m.bin"))
    open(scene, "w").write(text)
    for payload in (None, b"NOTAMESH" + b"\0" * 64, b"PBRTMESH" + np.array([0, 3, 3], "<u4").tobytes() + b"\0" * 8):
        if payload is not None:
            open(tmp_path / "m.bin", "wb").write(payload)
        with pytest.raises(pb.PbrtGpuError) as e:
            pb.HostScene(scene, ncores=1)
        assert "binary mesh" in str(e.value)
    pb.write_binary_mesh(str(tmp_path / "m.bin"), [0, 1, 2], [[0, 0, 0], [1, 0, 0], [0, 1, 0]])
    assert pb.HostScene(scene, ncores=1).counts()["triangles"] == pb.HostScene(scenes.variant("bunny", scene + ".ref.pbrt", 8, 8, 1), ncores=1).counts()["triangles"] + 1
