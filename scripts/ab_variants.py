#!/usr/bin/env python
"""A/B of kernel-variant libraries against the product library (run on the GPU box).

    python scripts/ab_variants.py [--workload bunny-c2] [--spp 128] [--steps 2] lib1.so lib2.so ...
    python scripts/ab_variants.py --env PBRTGPU_FOO=1 --env PBRTGPU_FOO=2        (same library, different switches)

Each arm is one `bench.py` run in its own process (PBRTGPU_LIB=<lib>, development spp) that also saves its film.  Whether a
variant computes the SAME image is decided on the pixels whose samples all come from their own render task
(weightSum == spp): those are order-independent in the reference; a pixel that received an edge splat from a neighbouring
task is summed in an order the reference itself leaves undefined (film/image.cpp:122-127), so a whole-film checksum
cannot tell a race from a bug.  Prints one JSON object.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def run_arm(lib, extra_env, workload, spp, steps, tmp, tag, timeout):
    env = dict(os.environ)
    env.setdefault("PBRTGPU_SCRATCH_GB", "24")
    if lib:
        env["PBRTGPU_LIB"] = lib
    for kv in extra_env:
        k, v = kv.split("=", 1)
        env[k] = v
    film = os.path.join(tmp, "film_%s.npy" % tag)
    cmd = [sys.executable, os.path.join(REPO, "bench.py"), "--workload", workload, "--steps", str(steps), "--warmup", "2",
           "--no-cpu-baseline", "--film-out", film]
    if spp:
        cmd += ["--spp", str(spp)]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout, env=env, cwd=REPO)
    rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    if not rows:
        raise RuntimeError((r.stderr or "no output").strip()[-400:])
    d = rows[-1]
    return {"value": d["value"], "e2e": d["e2e"]["value"], "kernel_ms": d.get("kernel_ms"),
            "roofline_frac": d["roofline"].get("frac"), "step_frac": d["roofline"].get("step_frac"),
            "k_trace_ms_per_step": d["roofline"].get("k_trace_ms_per_step"),
            "tail_ms_per_step": d["roofline"].get("k_path2_ms_per_step")}, np.load(film), d


def compare(film, ref, spp):
    own = (ref[..., 3] == spp) & (film[..., 3] == spp)
    same_w = bool(np.array_equal(film[..., 3], ref[..., 3]))
    a, b = film[..., :3][own].view(np.uint32), ref[..., :3][own].view(np.uint32)
    bad = int((a != b).any(axis=1).sum())
    return {"same_weights": same_w, "own_pixels": int(own.sum()), "own_pixels_differing": bad,
            "same_film_on_own_pixels": same_w and bad == 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("libs", nargs="*", help="variant libraries (paths relative to pbrt_v2_b200/, as PBRTGPU_LIB takes them)")
    ap.add_argument("--env", action="append", default=[], help="K=V switch for an extra arm of the product library")
    ap.add_argument("--workload", default="bunny-c2")
    ap.add_argument("--spp", type=int, default=128)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--timeout", type=float, default=300.0)
    args = ap.parse_args()
    import bench
    spp = args.spp or bench.WORKLOADS[args.workload][3]
    out = {"workload": args.workload, "spp": spp, "steps": args.steps, "arms": {}}
    with tempfile.TemporaryDirectory() as tmp:
        base, ref, _ = run_arm(None, [], args.workload, args.spp, args.steps, tmp, "product", args.timeout)
        out["arms"]["product"] = base
        arms = [(os.path.basename(l), l, []) for l in args.libs] + [(kv, None, [kv]) for kv in args.env]
        for i, (name, lib, ev) in enumerate(arms):
            try:
                v, film, _ = run_arm(lib, ev, args.workload, args.spp, args.steps, tmp, "v%d" % i, args.timeout)
                v.update(compare(film, ref, spp))
                v["speedup_vs_product"] = v["value"] / base["value"] if base["value"] else None
                out["arms"][name] = v
            except Exception as e:   # noqa: BLE001
                out["arms"][name] = {"error": repr(e)[:400]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
