"""The multi-GPU exchange step inside the host plugin: the library's NCCL communicator behind the C ABI
(pbrtgpu_comm_*, pbrtgpu_render_sharded) and the `pbrt_b200` CLI started once per GPU."""
import os
import subprocess

import numpy as np
import pytest

import pbrt_v2_b200 as pb
import oracle_api as oa
from helpers import host_scene

pytestmark = pytest.mark.gpu


def test_sharded_render_with_a_one_rank_communicator(staged):
    """world size 1: ncclGetUniqueId / ncclCommInitRank through the run-time loaded libnccl, then the sharded entry point
    must give pbrtgpu_render's film."""
    hs = host_scene("bunny", 48, 48, 8, 72)
    gs = pb.GpuScene(hs)
    want, _ = gs.render()
    comm = pb.Comm(pb.comm_unique_id(), 0, 1)
    got, st = gs.render_sharded(hs.render_desc(), comm)
    comm.close()
    assert st.cameraSamples == 48 * 48 * 8
    assert np.array_equal(got["Lxyz"].view(np.uint32), want["Lxyz"].view(np.uint32))
    assert np.array_equal(got["weightSum"], want["weightSum"])


@pytest.mark.skipif(pb.device_count() < 2, reason="needs two GPUs on the box (gpurun --gpus 2)")
def test_cli_on_two_gpus_writes_the_one_gpu_image(staged, tmp_path):
    """two `pbrt_b200` processes (rank 0 / 1, rendezvous file, one GPU each): rank 0's image must be the one-GPU image on
    every pixel no other task splats into, and the oracle's."""
    from pbrt_v2_b200 import scenes
    # (a Film "filename" in the scene wins over --outfile, as in the reference: film/image.cpp:258-269)
    path2 = scenes.variant("killeroo-simple", os.path.join(scenes.STAGED, "multi-cli-two.pbrt"), 64, 64, 8)
    path = scenes.variant("killeroo-simple", os.path.join(scenes.STAGED, "multi-cli-one.pbrt"), 64, 64, 8)
    exe = os.path.join(pb.REPO, "pbrt_v2_b200", "pbrt_b200")
    rdv = str(tmp_path / "nccl.id")
    procs = [subprocess.Popen([exe, "--ncores", "128", "--quiet", "--rank", str(r), "--world", "2", "--rendezvous", rdv,
                               os.path.basename(path2)], cwd=scenes.STAGED) for r in (0, 1)]
    assert all(p.wait(timeout=300) == 0 for p in procs)
    subprocess.run([exe, "--ncores", "128", "--quiet", os.path.basename(path)], cwd=scenes.STAGED, check=True, timeout=300)
    a, b = oa.read_pfm(os.path.join(scenes.STAGED, "multi-cli-two.pfm")), oa.read_pfm(os.path.join(scenes.STAGED, "multi-cli-one.pfm"))
    hs = pb.HostScene(path, ncores=128)
    ref, _ = oa.render(hs, threads=2)
    own = ref["weightSum"] == 8
    assert np.array_equal(a.view(np.uint32)[own], b.view(np.uint32)[own])
    assert np.array_equal(a.view(np.uint32)[own], hs.pixels_to_rgb(ref).view(np.uint32)[own])
