#!/usr/bin/env python
"""Golden vectors for InfiniteAreaLight's constructor, minted from the UNMODIFIED reference's own members
(oracle/_ref/bin/refdump envmap: radianceMap level 0, the Distribution2D's arrays; run where /root/reference exists):

  envmaps.npz   <name>_pfm / <name>_tga: the image file's bytes; <name>_L: L x scale; <name>_{head,data}: the dump
                (layout: pbrt_v2_b200.parse_envmap_dump)
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


def ref_envmap(pfm, L):
    """the reference's InfiniteAreaLight(identity, L, 1, pfm) -> (head uint32[4], data float32[])"""
    with tempfile.TemporaryDirectory() as tmp:
        fout = os.path.join(tmp, "o.bin")
        subprocess.run([REFDUMP, "envmap", pfm or "-"] + [repr(float(x)) for x in L] + [fout], check=True)
        raw = open(fout, "rb").read()
    return np.frombuffer(raw[:16], np.uint32).copy(), np.frombuffer(raw[16:], np.float32).copy()


def main():
    from helpers import ENV_MAPS, write_env_pfm
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for i, (fn, spec) in enumerate(sorted(ENV_MAPS.items())):
            path = os.path.join(tmp, fn)
            write_env_pfm(path, *spec)
            L = np.array([[1, 1, 1], [.5, .6, .9], [2, .25, 0]][i % 3], np.float32)
            key = fn[:-4].replace("-", "_")
            out[key + "_pfm"] = np.frombuffer(open(path, "rb").read(), np.uint8)
            out[key + "_L"] = L
            out[key + "_head"], out[key + "_data"] = ref_envmap(path, L)
        from helpers import TGA_MAPS, write_tga
        for fn, spec in sorted(TGA_MAPS.items()):
            path = os.path.join(tmp, fn)
            write_tga(path, *spec)
            L = np.array([1.5, 1, 2], np.float32)
            key = fn[:-4].replace("-", "_")
            out[key + "_tga"] = np.frombuffer(open(path, "rb").read(), np.uint8)
            out[key + "_L"] = L
            out[key + "_head"], out[key + "_data"] = ref_envmap(path, L)
        out["nomap_L"] = np.array([.4, .5, .7], np.float32)
        out["nomap_head"], out["nomap_data"] = ref_envmap(None, out["nomap_L"])
    np.savez_compressed(os.path.join(HERE, "envmaps.npz"), **out)
    print("wrote envmaps.npz: %d arrays" % len(out))


if __name__ == "__main__":
    main()
