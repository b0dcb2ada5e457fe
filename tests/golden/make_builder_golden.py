#!/usr/bin/env python
"""Golden vectors for the two scene-preparation builders, minted from the UNMODIFIED reference's own arrays
(oracle/_ref/bin/refdump bvh / loopsubdiv; run where /root/reference exists):

  builders.npz   bvh_<kind>_<n>_<maxp>_{bounds,nodes,order}: BVHAccel's LinearBVHNode[] and reordered primitive list over
                 box primitives; subdiv_<name>_L<levels>_{idx,P}: control mesh, _{rP,rN,ridx}: LoopSubdiv::Refine's TriangleMesh
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
REFDUMP = os.path.join(REPO, "oracle", "_ref", "bin", "refdump")

NODE_DUMP = np.dtype([("bmin", np.float32, 3), ("bmax", np.float32, 3), ("offset", np.uint32), ("nPrimitives", np.uint8),
                      ("axis", np.uint8), ("pad", np.uint8, 2)])


def ref_bvh(bounds, maxp):
    """the reference's BVHAccel over box primitives -> (nodes as NODE_DUMP records, order)"""
    bounds = np.ascontiguousarray(bounds, np.float32)
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "b.bin"), os.path.join(tmp, "o.bin")
        bounds.tofile(fin)
        subprocess.run([REFDUMP, "bvh", fin, str(maxp), fout], check=True)
        raw = open(fout, "rb").read()
    nn = int(np.frombuffer(raw, np.uint32, 1)[0])
    nodes = np.frombuffer(raw, NODE_DUMP, nn, 4)
    order = np.frombuffer(raw, np.uint32, len(bounds), 4 + 32 * nn)
    return nodes.copy(), order.copy()


def ref_loopsubdiv(indices, P, nlevels):
    """the reference's LoopSubdiv::Refine -> (P, N, indices) of the TriangleMesh it creates"""
    indices, P = np.ascontiguousarray(indices, np.int32).reshape(-1, 3), np.ascontiguousarray(P, np.float32).reshape(-1, 3)
    with tempfile.TemporaryDirectory() as tmp:
        fin, fout = os.path.join(tmp, "m.bin"), os.path.join(tmp, "o.bin")
        with open(fin, "wb") as f:
            f.write(np.array([len(indices), len(P)], np.int32).tobytes())
            f.write(indices.tobytes())
            f.write(P.tobytes())
        subprocess.run([REFDUMP, "loopsubdiv", fin, str(nlevels), fout], check=True)
        raw = open(fout, "rb").read()
    nv, nt = [int(x) for x in np.frombuffer(raw, np.uint32, 2)]
    rP = np.frombuffer(raw, np.float32, 3 * nv, 8).reshape(nv, 3)
    rN = np.frombuffer(raw, np.float32, 3 * nv, 8 + 12 * nv).reshape(nv, 3)
    ridx = np.frombuffer(raw, np.int32, 3 * nt, 8 + 24 * nv).reshape(nt, 3)
    return rP.copy(), rN.copy(), ridx.copy()


def main():
    from test_bvhbuild_on_host import boxes, KINDS
    from test_loopsubdiv_on_host import control_meshes
    out = {}
    rs = np.random.RandomState(2024)
    for kind in KINDS:
        for n, maxp in [(1, 1), (7, 1), (200, 4), (257, 255)]:
            b = boxes(kind, n, rs)
            nodes, order = ref_bvh(b, maxp)
            key = "bvh_%s_%d_%d" % (kind, n, maxp)
            out[key + "_bounds"], out[key + "_nodes"], out[key + "_order"] = b, nodes, order
    for name, (idx, P) in control_meshes():
        if name == "unused-vertices":
            continue   # (the reference's own constructor crashes on control meshes with unused vertices)
        for levels in (1, 2):
            try:
                rP, rN, ridx = ref_loopsubdiv(idx, P, levels)
            except subprocess.CalledProcessError as e:
                print("reference failed on", name, levels, e)
                continue
            key = "subdiv_%s_L%d" % (name, levels)
            out[key + "_idx"], out[key + "_P"] = np.asarray(idx, np.int32), np.asarray(P, np.float32)
            out[key + "_rP"], out[key + "_rN"], out[key + "_ridx"] = rP, rN, ridx
    np.savez_compressed(os.path.join(HERE, "builders.npz"), **out)
    print("wrote builders.npz: %d arrays" % len(out))


if __name__ == "__main__":
    main()
