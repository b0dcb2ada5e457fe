"""Scene-preparation kernels against the host front end: times pbrtgpu_bvh_build against the host front end's recursive build on the same input and checks that the trees are
the same (and pbrtgpu_loopsubdiv_refine against LoopSubdiv::Refine on the killeroo control mesh).  Inputs: the BASELINE scenes' own build inputs and random triangles' bounds at C5 size.  One JSON line per case.
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


def subdiv_case(nlevels, reps):
    """LoopSubdiv::Refine of the killeroo control mesh (scenes/geometry/killeroo.pbrt): device vs host front end"""
    import re
    from pbrt_v2_b200 import scenes
    text = open(os.path.join(scenes.STAGED, "geometry", "killeroo.pbrt")).read()
    m = re.search(r'Shape "loopsubdiv".*?"integer indices"\s*\[([^\]]*)\]\s*"point P"\s*\[([^\]]*)\]', text, re.S)
    cm = pb.ControlMesh(np.array(m.group(1).split(), np.int32).reshape(-1, 3), np.array(m.group(2).split(), np.float32).reshape(-1, 3))
    t0 = time.time()
    want = cm.host_refine(nlevels)
    host_s = time.time() - t0
    cm.refine(0)
    wall = []
    for _ in range(reps):
        t0 = time.time()
        got = cm.refine(nlevels)
        wall.append(time.time() - t0)
    same = bool(np.array_equal(got[2], want[2]) and np.array_equal(got[0], want[0], equal_nan=True) and
                np.array_equal(got[1], want[1], equal_nan=True))
    print(json.dumps({"case": "loopsubdiv killeroo nlevels %d" % nlevels, "control_faces": len(cm.indices), "triangles": len(want[2]),
                      "host_refine_s": round(host_s, 3), "device_call_wall_ms": round(1e3 * min(wall), 3),
                      "speedup_vs_host": round(host_s / min(wall), 1), "same_mesh": same}), flush=True)


def guard(name, fn):
    try:
        fn()
    except Exception as e:   # noqa: BLE001  (one failing case must not hide the others)
        print(json.dumps({"case": name, "error": repr(e)[:300]}), flush=True)


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
            guard(base, lambda: case(base, hs.build_input(), 4, a.reps))
        for nlevels in (1, 3):
            guard("loopsubdiv killeroo nlevels %d" % nlevels, lambda: subdiv_case(nlevels, a.reps))
    rs = np.random.RandomState(1)
    c = rs.rand(a.n, 3).astype(np.float32) * 2 - 1
    e = (rs.rand(a.n, 3).astype(np.float32) * 0.01).astype(np.float32)
    guard("random-%d" % a.n, lambda: case("random-%d" % a.n, np.concatenate([c - e, c + e], 1).astype(np.float32), 4, a.reps))


if __name__ == "__main__":
    main()
