# GPU vs oracle on the edge-case variants (tests/helpers.py EDGE_VARIANTS)
import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import pbrt_v2_b200 as pb
import helpers, oracle_api as oa
from pbrt_v2_b200 import scenes
scenes.stage()
for v in helpers.EDGE_VARIANTS:
    try:
        hs, _ = helpers.variant_scene(v[0])
        g, _ = pb.GpuScene(hs).render()
        c, _ = oa.render(hs, threads=2)
        same_w = np.array_equal(g["weightSum"], c["weightSum"])
        nd = int((g["Lxyz"].view(np.uint32) != c["Lxyz"].view(np.uint32)).any(axis=2).sum())
        print("EDGE", v[0], "weights", same_w, "differing pixels", nd, "of", g["weightSum"].size, flush=True)
    except Exception as e:
        print("EDGE", v[0], "ERR", repr(e)[:300], flush=True)
