#!/usr/bin/env python
"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container (needs /root/reference built into oracle/_ref by
`make -C oracle ref`):   python tests/golden/make_golden.py

  ld_samples.npz   LDSampler::GetMoreSamples tables (37 floats/sample) for a few
                   (spp, nTasks, taskNum), via oracle/_ref/bin/refdump
  rng.npz          MT19937 RandomUInt() prefixes for a few seeds
  images.npz       reference-rendered RGB images (PFM -> float32) of tiny variants of
                   the config scenes, at pinned nTasks (--ncores), + their scene settings
  counts.json      "BVH created with N nodes for M primitives" of the config scenes
"""
import json
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
from pbrt_v2_b200 import scenes
import oracle_api as oa

REFDUMP = os.path.join(REPO, "oracle", "_ref", "bin", "refdump")

LD_CASES = [  # (res, spp, nTasks, taskNum)
    (32, 4, 1024, 0), (32, 4, 1024, 37), (32, 64, 1024, 1023), (32, 1024, 64, 5), (48, 16, 4096, 131),
]
LDX_CASES = [  # (res, spp, nTasks, taskNum, n1D sizes, n2D sizes): DirectLightingIntegrator layouts (+ the 2 emission 1-D requests)
    (32, 4, 1024, 5, [8, 8, 1, 1], [8, 8]),            # killeroo-simple.pbrt as shipped: one area light, nsamples 8
    (32, 16, 64, 3, [1, 1, 4, 4, 1, 1], [1, 1, 4, 4]),  # bunny.pbrt under directlighting: point light + area light, nsamples 4
    (32, 64, 1024, 1000, [1, 1, 1, 1, 1], [1, 1]),      # strategy "one"
    (48, 1024, 64, 7, [16, 16, 1, 1], [16, 16]),
]
IMAGE_CASES = [  # (base, res, spp, ncores)
    ("bunny", 32, 4, 32), ("bunny", 32, 8, 1), ("killeroo-simple", 32, 4, 32), ("killeroo-simple", 32, 8, 1),
    # config C4's scene: two animated TransformedPrimitives, each over its own nested BVH
    ("anim-killeroos-moving", 32, 4, 32), ("anim-killeroos-moving", 32, 8, 1),
    # DirectLightingIntegrator (what the shipped scene files ask for): 5th field = integrator, 6th = strategy
    ("killeroo-simple", 32, 4, 32, "directlighting", None), ("killeroo-simple", 32, 8, 1, "directlighting", None),
    ("bunny", 32, 4, 32, "directlighting", None), ("bunny", 32, 8, 1, "directlighting", "one"),
    ("anim-killeroos-moving", 32, 4, 1, "directlighting", None),
]


def main():
    scenes.stage()
    ld = {}
    for res, spp, ntasks, task in LD_CASES:
        out = subprocess.run([REFDUMP, "ld", "0", str(res), "0", str(res), str(spp), str(res), str(res), str(ntasks),
                              str(task), "0", "1"], check=True, stdout=subprocess.PIPE, text=True).stdout.split("\n")
        n = int(out[0].split()[0])
        rows = np.array([[float(x) for x in out[1 + i].split()] for i in range(n)], dtype=np.float32)
        nxt = np.uint32(int(out[1 + n]))
        key = "r%d_s%d_n%d_t%d" % (res, spp, ntasks, task)
        ld[key] = rows
        ld[key + "_next"] = nxt
    np.savez_compressed(os.path.join(HERE, "ld_samples.npz"), **ld)
    ldx = {}
    for ci, (res, spp, ntasks, task, s1, s2) in enumerate(LDX_CASES):
        args = [REFDUMP, "ldx", "0", str(res), "0", str(res), str(spp), str(res), str(res), str(ntasks), str(task), "0", "1",
                str(len(s1))] + [str(x) for x in s1] + [str(len(s2))] + [str(x) for x in s2]
        out = subprocess.run(args, check=True, stdout=subprocess.PIPE, text=True).stdout.split("\n")
        n = int(out[0].split()[0])
        rows = np.array([[float(x) for x in out[1 + i].split()] for i in range(n)], dtype=np.float32)
        key = "case%d" % ci
        ldx[key] = rows
        ldx[key + "_meta"] = np.array([res, spp, ntasks, task, len(s1)] + s1 + [len(s2)] + s2, dtype=np.int64)
        ldx[key + "_next"] = np.uint32(int(out[1 + n]))
    np.savez_compressed(os.path.join(HERE, "ld_samples_general.npz"), **ldx)
    rng = {}
    for seed in (0, 1, 5489, 123456):
        out = subprocess.run([REFDUMP, "rng", str(seed), "2000"], check=True, stdout=subprocess.PIPE, text=True).stdout.split()
        rng["seed%d" % seed] = np.array([int(x) for x in out], dtype=np.uint32)
    np.savez_compressed(os.path.join(HERE, "rng.npz"), **rng)
    # the reference's own BxDF classes on random direction pairs (what tools/bsdftest.cpp exercises): refdump bxdf
    bx = {}
    for exponent, seed in ((40.0, 7), (1.0 / 0.15, 8), (20000.0, 9)):
        out = subprocess.run([REFDUMP, "bxdf", repr(exponent), "2000", str(seed)], check=True, stdout=subprocess.PIPE,
                             text=True).stdout.strip().split("\n")
        rows = np.array([[float.fromhex(x) for x in ln.split()] for ln in out], dtype=np.float32)
        bx["e%d" % seed] = rows
        bx["e%d_exponent" % seed] = np.float32(exponent)
    np.savez_compressed(os.path.join(HERE, "bxdf.npz"), **bx)
    imgs = {}
    have = os.path.join(HERE, "images.npz")
    if os.path.exists(have) and "--all" not in sys.argv:   # keep what is committed, mint only new cases
        imgs = dict(np.load(have))
    for case in IMAGE_CASES:
        base, res, spp, ncores = case[:4]
        integ = case[4] if len(case) > 4 else "path"
        strategy = case[5] if len(case) > 5 else None
        key = "%s_r%d_s%d_c%d" % (base.replace("-", "_"), res, spp, ncores)
        if integ != "path":
            key = "%s_%s_%s" % (integ if strategy is None else integ + strategy, "x", key)
        if key in imgs:
            continue
        p = scenes.variant(base, os.path.join(scenes.STAGED, "golden-%s.pbrt" % base), res, res, spp, integrator=integ,
                           strategy=strategy)
        rgb, _ = oa.run_reference(p, ncores)
        imgs[key] = rgb.astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "images.npz"), **imgs)
    counts = {}
    for base in ("bunny", "killeroo-simple"):
        # (short file name: the reference truncates its log lines at 80 columns)
        p = scenes.variant(base, os.path.join(scenes.STAGED, "g%s.pbrt" % base[0]), 8, 8, 1, filename="g.pfm")
        out = subprocess.run([oa.REF_BIN, "--ncores", "1", "--verbose", os.path.basename(p)], cwd=scenes.STAGED,
                             stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        m = re.search(r"BVH created with (\d+) nodes for (\d+)", out)
        counts[base] = {"nodes": int(m.group(1)), "primitives": int(m.group(2))}
    json.dump(counts, open(os.path.join(HERE, "counts.json"), "w"), indent=1, sort_keys=True)
    # Renderer "aggregatetest" (renderers/aggregatetest.cpp): the reference's accelerator check prints one
    # "Disagreement" warning per ray where BVHAccel and the exhaustive loop differ (rays that start on a flat
    # primitive with mint == 0); their number and the rays themselves are the known answer
    agg = {}
    for base in ("bunny", "killeroo-simple"):
        n = 20000
        p = scenes.variant(base, os.path.join(scenes.STAGED, "golden-agg-%s.pbrt" % base), 32, 32, 4)
        text = open(p).read().replace("WorldBegin", 'Renderer "aggregatetest" "integer niters" [%d]\nWorldBegin' % n, 1)
        open(p, "w").write(text)
        out = subprocess.run([oa.REF_BIN, "--ncores", "1", os.path.basename(p)], cwd=scenes.STAGED, stdout=subprocess.PIPE,
                             stderr=subprocess.STDOUT, text=True).stdout
        flat = re.sub(r"\s+", " ", out)
        rays = re.findall(r"Ray: org \[([^\]]+)\], dir \[([^\]]+)\], mint = (\S+)", flat)
        hx = lambda t: [float.fromhex(x.strip()) for x in t.split(",")]
        agg[base] = {"niters": n, "disagreements": len(rays),
                     "first_rays": [{"o": hx(a), "d": hx(b), "mint": float.fromhex(c)} for a, b, c in rays[:50]]}
    json.dump(agg, open(os.path.join(HERE, "aggregatetest.json"), "w"), indent=1, sort_keys=True)
    print("golden fixtures written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
