#!/usr/bin/env python
"""The scene front end (parser, shapes, Loop subdivision, host BVH build, flattening) under AddressSanitizer + UBSan on random
in-scope scenes and on mutated ones (characters flipped, tokens deleted / duplicated / swapped, numbers replaced by extreme
values, truncations).  CPU only.  usage: sanitize_parser.py [n] [seed]"""
import os, subprocess, sys, tempfile, glob, re
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "scripts"))
import fuzz_scenes

EXTREME = ["0", "-1", "1e30", "-1e30", "nan", "inf", "2147483647", "-2147483648", "4294967296", "1e-40", "999999999999", ""]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    sys.argv = sys.argv[:1] + ["0", "0", "specular+metal+crop+infinite+filter+animated+instances+params"]   # what fuzz_scenes.gen reads its flags from
    tmp = tempfile.mkdtemp(prefix="pbrt_san_parser")
    exe = os.path.join(tmp, "san_parser")
    host = os.path.join(REPO, "pbrt_v2_b200", "csrc", "host")
    srcs = [f for f in glob.glob(os.path.join(host, "*.cpp")) if os.path.basename(f) not in ("main.cpp", "capi.cpp")]
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
                    "-I" + os.path.join(REPO, "include"), os.path.join(REPO, "tests/native/san/main_parser.cpp")] + srcs +
                   ["-lz", "-lpthread", "-o", exe], check=True)
    os.makedirs(os.path.join(tmp, "brdfs"), exist_ok=True)
    from pbrt_v2_b200 import scenes
    brdf = os.path.join(scenes.STAGED, "brdfs", "mystique.brdf")
    if os.path.exists(brdf):
        os.symlink(brdf, os.path.join(tmp, "brdfs", "mystique.brdf"))
    rs = np.random.RandomState(seed)
    seeds = []
    for k in range(40):
        p = os.path.join(tmp, "seed%02d.pbrt" % k)
        fuzz_scenes.gen(np.random.RandomState(900000 + 1000 * seed + k), p)
        seeds.append(p)
    r = subprocess.run([exe] + seeds, capture_output=True, text=True)
    print(r.stdout.strip())
    if r.returncode != 0:
        print(r.stderr[-3000:]); sys.exit(1)
    batch, done = [], 0
    for k in range(n):
        text = open(seeds[k % len(seeds)]).read()
        kind = rs.choice([0, 1, 1, 1, 2, 3, 4, 5])
        if kind == 0:      # flip characters
            b = bytearray(text.encode())
            for _ in range(rs.randint(1, 4)): b[rs.randint(len(b))] = rs.randint(32, 127)
            text = b.decode("latin1")
        elif kind == 1:    # numbers -> extreme values
            nums = list(re.finditer(r"-?\d+\.?\d*(e-?\d+)?", text))
            for _ in range(rs.randint(1, 4)):
                m = nums[rs.randint(len(nums))]
                text = text[:m.start()] + EXTREME[rs.randint(len(EXTREME))] + text[m.end():]
                nums = list(re.finditer(r"-?\d+\.?\d*(e-?\d+)?", text))
                if not nums: break
        elif kind == 2:    # truncate
            text = text[:rs.randint(len(text))]
        else:              # delete / duplicate / swap tokens or lines
            unit = text.split("\n") if kind == 3 else text.split(" ")
            for _ in range(rs.randint(1, 4)):
                if len(unit) < 3: break
                i, j = rs.randint(len(unit)), rs.randint(len(unit))
                op = rs.randint(3)
                if op == 0: del unit[i]
                elif op == 1: unit.insert(j, unit[i])
                else: unit[i], unit[j] = unit[j], unit[i]
            text = ("\n" if kind == 3 else " ").join(unit)
        p = os.path.join(tmp, "m%05d.pbrt" % k)
        open(p, "w", encoding="latin1").write(text); batch.append(p)
        if len(batch) == 200 or k == n - 1:
            try:
                r = subprocess.run([exe] + batch, capture_output=True, text=True, errors="replace", timeout=900)
            except subprocess.TimeoutExpired:
                print("TIMEOUT in batch ending at", k, "(kept under", tmp, ")"); sys.exit(1)
            if r.returncode != 0:
                print(r.stderr[-4000:]); print("FAILED in batch ending at", k, "(kept under", tmp, ")"); sys.exit(1)
            done += len(batch)
            for f in batch: os.remove(f)
            batch = []
    print("mutated scenes loaded or refused without a sanitizer report:", done)


if __name__ == "__main__":
    main()
