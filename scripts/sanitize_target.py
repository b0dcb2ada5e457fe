#!/usr/bin/env python
"""Small renders of every kernel family for compute-sanitizer (scripts/gpu_sanitize.sh): both integrators, a static
scene with the measured BRDF, the animated scene (TransformedPrimitives), specular materials, multi-pixel tasks (rounds),
ray-batch entry points, LD-table test hook, device BVH build and LoopSubdiv."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import pbrt_v2_b200 as pb                      # noqa: E402
from pbrt_v2_b200 import scenes                # noqa: E402
from helpers import host_scene, variant_scene, random_rays, scene_bounds   # noqa: E402

scenes.stage()
res = int(os.environ.get("SANITIZE_RES", "24"))
cases = [("bunny", res, 8, 32, "path"), ("bunny", res, 8, 1, "path"), ("killeroo-simple", res, 4, 2, "directlighting"),
         ("anim-killeroos-moving", res, 4, 2, "path"), ("killeroo-simple", res, 64, 32, "path")]
for base, r, spp, ncores, integ in cases:
    hs = host_scene(base, r, r, spp, ncores, integrator=integ)
    gs = pb.GpuScene(hs)
    px, st = gs.render()
    print(base, integ, spp, ncores, "samples", st.cameraSamples, "sum", float(px["Lxyz"].sum()))
    if base == "bunny" and ncores == 32:
        lo, hi = scene_bounds(hs)
        rays = random_rays(20000, 3, lo, hi)
        print("  intersect", int(gs.intersect(rays)["hit"].sum()), "occluded", int(gs.intersect_p(rays).sum()))
        rd = hs.render_desc()
        for t in range(rd.nTasks):   # the first task with a non-empty pixel window
            try:
                print("  ld_samples of task", t, float(pb.ld_samples(rd, t).sum()))
                break
            except pb.PbrtGpuError:
                continue
    gs.close()
hs, _ = variant_scene("path-glass-mirror")
px, st = pb.GpuScene(hs).render()
print("path-glass-mirror", float(px["Lxyz"].sum()))
b = np.random.RandomState(1).rand(5000, 3).astype(np.float32)
bounds = np.concatenate([b, b + 0.01], axis=1)
nodes, order, bst = pb.bvh_build(bounds)
print("bvh_build", len(nodes), bst.buildMs)
print("ok")
