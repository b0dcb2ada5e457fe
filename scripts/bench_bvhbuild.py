"""Times pbrtgpu_bvh_build against the host front end's recursive build on the same input and checks that the trees are
the same.  Inputs: the BASELINE scenes' own build inputs and random triangles' bounds at C5 size.  One JSON line per case.
Run on a GPU box:  python scripts/bench_bvhbuild.py [--n 10000000]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import pbrt_v2_b200 as pb   # noqa: E402


def same_tree(a, b):
    return len(a) == len(b) and all(np.array_equal(a[k], b[k]) for k in ("bmin", "bmax", "offset", "nPrimitives", "axis"))


def case(name, bounds, maxp, reps):
    t0 = time.time()
    want_nodes, want_order = pb.host_bvh_build(bounds, maxp)
    host_s = time.time() - t0
    pb.bvh_build(bounds[:64], maxp)   # context creation outside the timing
    ms, wall = [], []
    for _ in range(reps):
        t0 = time.time()
        nodes, order, st = pb.bvh_build(bounds, maxp)
        wall.append(time.time() - t0)
        ms.append(st.buildMs)
    print(json.dumps({"case": name, "primitives": len(bounds), "maxPrimsInNode": maxp, "nodes": len(nodes), "levels": st.nLevels,
                      "host_build_s": round(host_s, 3), "device_build_ms": round(min(ms), 3),
                      "device_call_wall_ms": round(1e3 * min(wall), 3),   # + the NaN scan and result copies on the host side
                      "Mprims_per_s_device": round(len(bounds) / min(ms) / 1e3, 1),
                      "speedup_vs_host": round(host_s * 1e3 / min(ms), 1),
                      "same_tree": bool(np.array_equal(order, want_order) and same_tree(nodes, want_nodes))}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10000000, help="random-triangle case size (BASELINE C5: 10 M)")
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    from helpers import host_scene
    from pbrt_v2_b200 import scenes
    scenes.stage()
    if scenes.have_staged():
        for base in ("bunny", "killeroo-simple"):
            pb.keep_build_input(True)
            hs = host_scene(base, 32, 32, 4, 32)
            pb.keep_build_input(False)
            case(base, hs.build_input(), 4, a.reps)
    rs = np.random.RandomState(1)
    c = rs.rand(a.n, 3).astype(np.float32) * 2 - 1
    e = (rs.rand(a.n, 3).astype(np.float32) * 0.01).astype(np.float32)
    case("random-%d" % a.n, np.concatenate([c - e, c + e], 1).astype(np.float32), 4, a.reps)


if __name__ == "__main__":
    main()
