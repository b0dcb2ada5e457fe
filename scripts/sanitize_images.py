#!/usr/bin/env python
"""The image readers of the host front end (.pfm / .tga / .exr) under AddressSanitizer + UBSan on valid files and on
mutated ones (bit flips, overwritten header fields, truncations, garbage tails).  CPU only.  usage: sanitize_images.py [n]"""
import os, subprocess, sys, tempfile
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
from helpers import write_env_pfm, write_tga, ENV_MAPS, TGA_MAPS
from test_exr import write_exr, _image


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    tmp = tempfile.mkdtemp(prefix="pbrt_san_img")
    exe = os.path.join(tmp, "san_images")
    host = os.path.join(REPO, "pbrt_v2_b200", "csrc", "host")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                    "-I" + os.path.join(REPO, "include"), os.path.join(REPO, "tests/native/san/main_images.cpp")] +
                   [os.path.join(host, f) for f in ("envmap.cpp", "tga.cpp", "exr.cpp", "measured.cpp")] + ["-lz", "-o", exe], check=True)
    seeds = []
    for fn, spec in ENV_MAPS.items():
        write_env_pfm(os.path.join(tmp, fn), *spec); seeds.append(os.path.join(tmp, fn))
    for fn, spec in TGA_MAPS.items():
        write_tga(os.path.join(tmp, fn), *spec); seeds.append(os.path.join(tmp, fn))
    rs = np.random.RandomState(0)
    for c in range(4):
        img = _image(rs, 9 + 8 * c, 7 + c)
        p = os.path.join(tmp, "s%d.exr" % c)
        ch = {"R": (1, img[..., 0]), "G": (1, img[..., 1]), "B": (1, img[..., 2])}
        if c & 1: ch["A"] = (2, np.ones(img.shape[:2], np.float32))
        write_exr(p, ch, c); seeds.append(p)
    from pbrt_v2_b200 import scenes
    brdf = os.path.join(scenes.STAGED, "brdfs", "mystique.brdf")
    if os.path.exists(brdf):   # the measured-BRDF reader (a text file of floats, materials/measured.cpp:85-135): the first 400 records
        tok = [t for line in open(brdf) for t in line.split("#")[0].split()][:20000]
        nw = int(float(tok[0]))
        keep = tok[:1 + nw + 400 * (4 + nw)]
        open(os.path.join(tmp, "s.brdf"), "w").write("\n".join(keep) + "\n")
        seeds.append(os.path.join(tmp, "s.brdf"))
    subprocess.run([exe] + seeds, check=True)
    batch, done = [], 0
    for k in range(n):
        src = seeds[k % len(seeds)]
        d = bytearray(open(src, "rb").read())
        kind = rs.randint(5)
        if kind == 0:
            for _ in range(rs.randint(1, 4)): d[rs.randint(len(d))] ^= 1 << rs.randint(8)
        elif kind == 1:
            at = rs.randint(min(len(d), 96)); d[at] = rs.randint(256)
        elif kind == 2:
            d = d[:rs.randint(len(d))]
        elif kind == 3:
            at = rs.randint(min(len(d) - 4, 400)); d[at:at + 4] = rs.randint(0, 256, 4).astype(np.uint8).tobytes()
        else:
            at = rs.randint(len(d)); d[at:] = rs.randint(0, 256, len(d) - at).astype(np.uint8).tobytes()
        p = os.path.join(tmp, "m%05d%s" % (k, os.path.splitext(src)[1]))
        open(p, "wb").write(bytes(d)); batch.append(p)
        if len(batch) == 250 or k == n - 1:
            r = subprocess.run([exe] + batch, capture_output=True, text=True)
            if r.returncode != 0:
                print(r.stderr[-3000:]); print("FAILED in batch ending at", k); sys.exit(1)
            done += len(batch)
            for f in batch: os.remove(f)
            batch = []
    print("mutated files read without a sanitizer report:", done)


if __name__ == "__main__":
    main()
