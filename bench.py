#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 path-tracing hot path.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference arm: CPU)

Metric (BASELINE.json): Msamples/s = camera samples (W*H*spp) per second of render
phase.  Workload at every N: BASELINE.json configs[1] -- scenes/bunny.pbrt, path
integrator, 1024 spp, 1024x1024, BVH accelerator -- one "step" = one full
Renderer::Render of that image.  For N > 1 the image's render tasks are sharded
across ranks (interleaved blocks of rows, no data-path collective) and the
framebuffer is summed to rank 0 with one NCCL reduce; total work is fixed, so
"scaling" is "strong".

Prints ONE JSON line on rank 0 (see the keys below).  Numbers are device-timed
(CUDA events on the launching stream), max over ranks.  Side measurements (kernel
variants, scene-preparation kernels) are NOT part of this command: scripts/ab_variants.py.
"""
import argparse
import ctypes as C
import json
import os
import re
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (staged base scene, xres, yres, spp)
    "bunny-c2": ("bunny", 1024, 1024, 1024),            # BASELINE.json configs[1]
    "killeroo-c1": ("killeroo-simple", 400, 400, 64),   # configs[0] (CPU-runnable case)
    "killeroo-c3": ("killeroo-simple", 1920, 1080, 4096),
    # configs[3]: two animated TransformedPrimitives (each its own nested BVH) of killeroos subdivided to
    # 532 224 triangles (geometry copy with "integer nlevels" [3]), shutter 0..1
    "anim-c4": ("anim-killeroos-moving", 1920, 1080, 1024),
    # scenes/killeroo-simple.pbrt exactly as shipped (SurfaceIntegrator "directlighting", strategy
    # "all", area light with 8 samples: 17 rays per camera sample), 700x700, 64 spp
    "killeroo-dl": ("killeroo-simple", 700, 700, 64),
    # configs[4]: synthetic 10M-triangle random mesh, maxdepth 8 (HBM-resident: 1.3 GB of BVH + triangles)
    "synthetic-c5": ("synthetic", 4096, 4096, 512),
}
WORKLOAD_NLEVELS = {"anim-c4": 3}
WORKLOAD_INTEGRATOR = {"killeroo-dl": "directlighting"}
SYNTHETIC_TRIS = 10000000
# dram__bytes_read.sum + dram__bytes_write.sum of one mid-render k_wtrace launch (ncu --set full, profiles/), per workload
K_TRACE_DRAM_BYTES_PER_LAUNCH = {
    "bunny-c2": 635614976 + 241986048,   # profiles/r02g_k_wtrace_ncu_full.txt (launch 60 of a 256-spp run: a full 4 Mi-slot pool, as at 1024 spp)
}
CPU_SAMPLE_SPP = 16     # bounded sample of the same workload for the CPU arms (cost is linear in spp)


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def per_pixel_ncores(xres, yres):
    """--ncores value that makes the reference create the finest task grid it can: one render task per pixel,
    nTasks = RoundUpPow2(max(32*ncores, nPixels/256)) >= nPixels (samplerrenderer.cpp:203-205), unless that
    overflows the int products of Sampler::ComputeSubWindow (sampler.cpp:58, 2*dx*ny < dy*nx) -- then the
    largest power of two that does not (4096x4096: 2^18 tasks of 8x8 pixels)."""
    npx = xres * yres
    n = 1
    while n < npx:
        n *= 2
    def overflows(count):
        nx, ny = count, 1
        while True:
            if 2 * xres * ny > 0x7fffffff or yres * nx > 0x7fffffff:
                return True
            if not (nx % 2 == 0 and 2 * xres * ny < yres * nx):
                return False
            nx //= 2
            ny *= 2
    while n > 32 and overflows(n):
        n //= 2
    return max(1, n // 32)


def tasks_note(xres, yres):
    n = 32 * per_pixel_ncores(xres, yres)
    if n >= xres * yres:
        return "one reference render task per pixel (--ncores %d)" % (n // 32)
    return "%d reference render tasks of ~%.3g pixels (--ncores %d; finer grids overflow Sampler::ComputeSubWindow's int products)" % (
        n, xres * yres / n, n // 32)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Threads the CPU arms may really use: the affinity mask, capped by the cgroup CPU quota (a leased box can show 128
    CPUs in os.cpu_count() and still be throttled to a fraction of them: BENCH_r01 vs SCALE_r01 differed 5.8x on the same
    binary).  Returns (threads, detail)."""
    n_aff = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    quota = None
    try:   # cgroup v2
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = float(q) / float(p)
    except Exception:
        try:   # cgroup v1
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = q / p
        except Exception:
            pass
    n = n_aff
    if quota is not None:
        n = max(1, min(n_aff, int(quota + 0.5)))
    return n, {"os_cpu_count": os.cpu_count(), "sched_affinity": n_aff, "cgroup_cpu_quota": quota, "threads_used": n}


def reference_render_seconds(scene_path, ncores):
    """Run the unmodified reference (oracle/_ref/bin/pbrt) and return its own render-phase time
    (ProgressReporter's final '(N.Ns)', core/progressreporter.cpp:60-118) and the wall time."""
    ref_bin = os.path.join(REPO, "oracle", "_ref", "bin", "pbrt")
    cwd = os.path.dirname(scene_path)
    t0 = time.time()
    p = subprocess.run([ref_bin, "--ncores", str(ncores), os.path.basename(scene_path)], cwd=cwd,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True)
    wall = time.time() - t0
    m = re.findall(r"\((\d+\.\d+)s\)", p.stdout)
    return (float(m[-1]) if m else wall), wall


def cpu_reference_rate(workload, threads):
    """Msamples/s of the reference's multithreaded CPU path on a bounded sample of `workload`."""
    from pbrt_v2_b200 import scenes
    base, xres, yres, _ = WORKLOADS[workload]
    path = scenes.variant(base, os.path.join(scenes.STAGED, "bench-cpu-%s.pbrt" % workload), xres, yres, CPU_SAMPLE_SPP,
                          nlevels=WORKLOAD_NLEVELS.get(workload), integrator=WORKLOAD_INTEGRATOR.get(workload, "path"))
    sec, wall = reference_render_seconds(path, threads)
    nsamp = xres * yres * CPU_SAMPLE_SPP
    return nsamp / max(sec, 1e-3) / 1e6, sec, wall


def run_reference_arm(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, all host cores."""
    if rank != 0:
        return
    workload = args.workload
    base, xres, yres, spp = WORKLOADS[workload]
    threads, thread_detail = host_threads()
    ref_bin = os.path.join(REPO, "oracle", "_ref", "bin", "pbrt")
    from pbrt_v2_b200 import scenes
    scenes.stage()
    if not os.path.exists(ref_bin) or not scenes.have_staged():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bin/pbrt or staged scenes missing (build() must run where /root/reference exists)"}))
        return
    for _ in range(args.warmup):
        cpu_reference_rate(workload, threads)
    secs = []
    for _ in range(args.steps):
        _, sec, _ = cpu_reference_rate(workload, threads)
        secs.append(sec)
    nsamp = xres * yres * CPU_SAMPLE_SPP
    total = sum(secs)
    value = nsamp * len(secs) / total / 1e6
    sample = "%s at %d spp (1/%d of the %d-spp workload; cost is linear in spp), %dx%d, --ncores %d" % (
        base, CPU_SAMPLE_SPP, spp // CPU_SAMPLE_SPP, spp, xres, yres, threads)
    line = {
        "impl": "reference", "metric": "Msamples/sec", "value": value, "unit": "Msamples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(secs),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the config names what this arm RAN: the bounded sample (CPU_SAMPLE_SPP of the workload's spp), default nTasks
        "config": workload_config(workload, world, per_pixel=False, spp_override=CPU_SAMPLE_SPP, sample_of=spp),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": threads, "kind": "reference", "sample": sample,
                         "host": thread_detail},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(workload, world, per_pixel=True, spp_override=None, sample_of=None):
    base, xres, yres, spp = WORKLOADS[workload]
    if spp_override:
        spp = spp_override
    desc = ("scenes/%s.pbrt%s, SurfaceIntegrator \"%s\" (maxdepth 5)" % (
        base, (" with geometry/killeroo.pbrt at \"integer nlevels\" [%d]" % WORKLOAD_NLEVELS[workload]) if workload in WORKLOAD_NLEVELS else "",
        WORKLOAD_INTEGRATOR.get(workload, "path"))) \
        if base != "synthetic" else \
        "synthetic random-triangle scene (SURVEY 8d C5), SurfaceIntegrator \"path\" (maxdepth 8)"
    sample_note = (" (a %d-spp sample of the %d-spp workload: cost is linear in spp)" % (spp, sample_of)) if sample_of else ""
    return {
        "workload": "%s, Sampler \"lowdiscrepancy\" %d spp%s, %dx%d, Accelerator \"bvh\", box filter" % (
            desc, spp, sample_note, xres, yres),
        "global_batch": xres * yres * spp, "parallelism": "tile%d" % world,
        "tasks": tasks_note(xres, yres) if per_pixel
                 else "reference default nTasks",
        "l2": "256 MiB buffer rewritten between steps (L2 flush); the per-step sample tables (64 B per camera sample, 69 GB at C2) stream through HBM",
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="bunny-c2", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override spp (development only; invalidates the number)")
    ap.add_argument("--tris", type=int, default=SYNTHETIC_TRIS, help="triangle count of the synthetic workload")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="(accepted for old scripts; side measurements now live in scripts/ab_variants.py)")
    ap.add_argument("--film-out", default="", help="save the last film as .npy (A/B runs: scripts/ab_variants.py compares "
                                                   "the pixels no other task splats into, bit for bit)")
    ap.add_argument("--economy", action="store_true",
                    help="large configs: the counting pass and the per-launch timing pass stand in for the warm-up steps")
    ap.add_argument("--lean", action="store_true",
                    help="(with --economy) no per-launch timing pass and no end-to-end region: the counting pass is the warm-up; "
                         "the line carries value + step_frac only and says so")
    args = ap.parse_args()

    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import pbrt_v2_b200 as pb
    from pbrt_v2_b200 import scenes

    if world > 1:
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    rc = pb.gpu_lib().pbrtgpu_set_device(local_rank)
    assert rc == 0, pb.gpu_lib().pbrtgpu_last_error()
    dev = torch.device("cuda", local_rank)

    scenes.stage()
    if not scenes.have_staged():
        raise SystemExit("fixture scenes missing: run __graft_entry__.build() where /root/reference exists")
    base, xres, yres, spp = WORKLOADS[args.workload]
    if args.spp:
        spp = args.spp
    t_prep = time.perf_counter()
    if base == "synthetic":
        hs = pb.HostScene.synthetic(args.tris, xres, yres, spp, maxdepth=8, ncores=per_pixel_ncores(xres, yres))
    else:
        path = scenes.variant(base, os.path.join(scenes.STAGED, "bench-%s-r%d.pbrt" % (args.workload, rank)), xres, yres, spp,
                              nlevels=WORKLOAD_NLEVELS.get(args.workload),
                              integrator=WORKLOAD_INTEGRATOR.get(args.workload, "path"))
        hs = pb.HostScene(path, ncores=per_pixel_ncores(xres, yres))
    prep_host_s = time.perf_counter() - t_prep   # parse + Loop subdivision + BVH build + flattening, on the host
    rd = hs.render_desc()
    from pbrt_v2_b200.multigpu import shard_render_desc
    shard_render_desc(rd, rank, world, xres)   # interleave blocks of 4 pixel rows across ranks
    comm = None
    if world > 1:   # the library's own NCCL communicator (pbrtgpu_comm_*): the exchange step lives in the host plugin
        ids = [pb.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = pb.Comm(ids[0], rank, world)
    _, _, xc, yc = hs.film_extent()
    nsamples = xc * yc * spp
    scene_bytes = hs.scene_bytes()

    t_up = time.perf_counter()
    gs = pb.GpuScene(hs)             # pbrtgpu_scene_create: validation + the host -> device copy of the flattened scene
    torch.cuda.synchronize(dev)
    prep_upload_s = time.perf_counter() - t_up
    film = torch.zeros((yc, xc, 4), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    host_film = torch.empty((yc, xc, 4), dtype=torch.float32, pin_memory=True)
    stream = torch.cuda.current_stream(dev)

    def step_resident(want_stats=False):
        if world == 1:
            return gs.render_device(rd, film.data_ptr(), stream.cuda_stream, want_stats=want_stats)
        _, st = gs.render_sharded(rd, comm, fetch=False)   # shard + render + ncclReduce to rank 0, film stays on the device
        return st

    g2 = pb.GpuScene(hs)   # the end-to-end arm's own handle: every step re-uploads the scene into it

    host_px = host_film.numpy().view(pb.PIXEL_DTYPE).reshape(yc, xc)   # the caller's (pinned) HOST pixel buffer

    def step_e2e():
        g2.update(hs)                              # host -> device copy of the flattened scene (the step's input)
        if world == 1:
            g2.render(rd, out=host_px)             # pbrtgpu_render: the host-buffer entry point (film D2H inside)
        else:
            g2.render_sharded(rd, comm, out=host_px if rank == 0 else None)   # + ncclReduce + rank 0's film D2H inside
        stream.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- untimed counting pass (ray / node / primitive-test counters) + warm-up
    gs.set_stats(1)
    counted = step_resident(want_stats=True)
    gs.set_stats(0)
    counts = torch.tensor([counted.closestRays, counted.shadowRays, counted.nodesVisited, counted.primTests],
                          dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    closest, shadow, nodes, prims = [float(x) for x in counts.tolist()]
    # ---- untimed roofline pass: the traversal kernel timed launch by launch (CUDA events on its stream) with the
    # device to itself (one lane), + the tail kernel's exclusive time
    if args.lean:
        rl_t = torch.zeros(4, dtype=torch.float64, device=dev)
    else:
        gs.set_stats(2)
        rl = step_resident(want_stats=True)
        gs.set_stats(0)
        rl_t = torch.tensor([rl.rayKernelMs, rl.tailMs, float(rl.rayKernelLaunches), rl.renderMs], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(rl_t, op=dist.ReduceOp.MAX)
    ray_ms, tail_excl_ms, ray_launches, serial_ms = [float(x) for x in rl_t.tolist()]
    # (the two passes above are full steps of the same work: with --economy they stand in for two warm-up steps)
    for _ in range(max(0, args.warmup - (2 if args.economy else 1))):
        flush.fill_(1)
        step_resident()

    # ---- timed: K steps, device-resident inputs
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    barrier()
    path_ms, table_ms, tail_ms, launches = 0.0, 0.0, 0.0, 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms = 0.0
    for _ in range(args.steps):
        flush.fill_(1)                              # L2 flush between timed iterations (not timed)
        e0.record(stream)
        st = step_resident(want_stats=True)
        e1.record(stream)
        e1.synchronize()
        total_ms += e0.elapsed_time(e1)
        path_ms += st.traceMs
        tail_ms += st.tailMs
        table_ms += st.tableMs
        launches += int(st.kernelLaunches)
    barrier()
    t = torch.tensor([total_ms, path_ms, table_ms, tail_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, path_ms, table_ms, tail_ms = [float(x) for x in t.tolist()]
    sampler.stop()
    clocks = sampler.summary()

    # ---- timed: K steps end to end (scene H2D + render + reduce + film D2H)
    if not args.economy:
        step_e2e()
    barrier()
    e2e_ms = 0.0
    for _ in range(0 if args.lean else args.steps):
        flush.fill_(1)
        e0.record(stream)
        step_e2e()
        e1.record(stream)
        e1.synchronize()
        e2e_ms += e0.elapsed_time(e1)
    barrier()
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())

    if rank == 0:
        value = nsamples * args.steps / (total_ms * 1e-3) / 1e6
        e2e_value = nsamples * args.steps / (e2e_ms * 1e-3) / 1e6 if e2e_ms > 0 else None
        # algorithmic bytes: 32 B per BVH node whose bounds are tested + 48 B per
        # primitive test (+36 B per shaded hit on a mesh with normals: none in bunny)
        alg_bytes = 32.0 * nodes + 48.0 * prims
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # BVH traversal happens in k_wtrace (phase 1, the vertex wavefront) and inside k_vertex<2> (phase 2, the RNG-driven
        # task chains); the node / primitive-test counters cover both, so the time does too: k_wtrace's launches timed one
        # by one in the roofline pass + the phase-2 kernel's exclusive time in that pass (it also shades: conservative)
        # Per GPU: the counters are summed over ranks, the times are one rank's (max over ranks), the peak is one GPU's.
        trace_ms = ray_ms + tail_excl_ms
        alg_bytes_gpu = alg_bytes / world
        achieved = alg_bytes_gpu / (trace_ms * 1e-3) / 1e9 if trace_ms > 0 else 0.0
        step_achieved = alg_bytes_gpu / (total_ms / args.steps * 1e-3) / 1e9
        hbm_resident = scene_bytes > 100e6   # B200 L2: 126 MB
        line = {
            "metric": "Msamples/sec", "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.workload, world, spp_override=spp),
            "e2e": {"value": e2e_value, "unit": "Msamples/s",
                    "h2d_bytes_per_step": scene_bytes, "d2h_bytes_per_step": int(xc * yc * 16),
                    "through": "pbrtgpu_scene_update + pbrtgpu_render (host pixel buffer)" if world == 1 else
                               "pbrtgpu_scene_update + pbrtgpu_render_sharded (ncclReduce inside the library, host pixel buffer on rank 0)"},
            "gpu_launches": launches,
            # outside the timed region, once per scene: what the step's input cost to prepare (rank 0's wall clock)
            "scene_prep": {"host_s": prep_host_s, "scene_create_s": prep_upload_s, "scene_bytes": scene_bytes,
                           "what": "host: parse + LoopSubdiv::Refine + BVHAccel build + flattening; scene_create: validation + upload"},
            "clocks": clocks,
            "roofline": {
                "bound": "hbm", "kernel": "k_wtrace (+ the traversals inside k_vertex<2>)", "achieved": achieved if trace_ms > 0 else None, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak if (peak and trace_ms > 0) else None,
                "per": "GPU (algorithmic bytes of all ranks / n_gpus, over one rank's kernel time)",
                # the same bytes over the whole timed step (tables, shading, film included): what the throughput implies
                "step_achieved": step_achieved, "step_frac": step_achieved / peak if peak else None,
                "traffic": K_TRACE_DRAM_BYTES_PER_LAUNCH.get(args.workload),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_sample": alg_bytes / nsamples,
                "algorithmic_bytes_per_launch": alg_bytes_gpu / max(1.0, ray_launches),
                "kernel_ms_per_launch": ray_ms / max(1.0, ray_launches),
                "launches_per_step": ray_launches,
                "k_trace_ms_per_step": ray_ms, "k_path2_ms_per_step": tail_excl_ms, "serial_step_ms": serial_ms,
                "k_trace_share_of_serial_step": ray_ms / serial_ms if serial_ms else None,
                "how": "untimed extra pass on one lane: cudaEventRecord around every k_wtrace launch on its stream "
                       "(pbrtgpu_set_stats(scene, 2)); in the timed steps the lanes' kernels overlap, so per-kernel "
                       "durations are not separable there",
                "note": ("scene (BVH+triangles, %.1f MB) exceeds the 126 MB L2: traversal streams from HBM" if hbm_resident else
                         "scene (BVH+triangles, %.1f MB) is L2-resident: the algorithmic-bytes roofline is an upper bound "
                         "on useful traffic, not DRAM traffic") % (scene_bytes / 1e6),
            },
            "rays": {"closest_per_sample": closest / nsamples, "shadow_per_sample": shadow / nsamples,
                     "nodes_per_ray": nodes / max(1.0, closest + shadow), "prim_tests_per_ray": prims / max(1.0, closest + shadow)},
            # per-lane event times summed over the concurrent lanes (they overlap: the sum exceeds ms_per_step)
            "kernel_ms": {"k_path_phase1": path_ms / args.steps, "k_path_phase2": tail_ms / args.steps,
                          "k_tables": table_ms / args.steps},
            "lanes": env_int("PBRTGPU_LANES", 2),
        }
        if spp != WORKLOADS[args.workload][3]:
            line["INVALID"] = "spp overridden to %d (development run)" % spp
        if args.lean:
            line["lean"] = ("--lean run (GPU budget): no per-launch timing pass (roofline.frac / kernel times absent, step_frac stands) "
                            "and no end-to-end region (e2e.value null); warm-up = the counting pass")
        if world == 1 and not args.no_cpu_baseline and base != "synthetic" and os.path.exists(os.path.join(REPO, "oracle", "_ref", "bin", "pbrt")):
            threads, thread_detail = host_threads()
            v, sec, wall = cpu_reference_rate(args.workload, threads)
            line["cpu_baseline"] = {
                "value": v, "unit": "Msamples/s", "cores": threads, "kind": "reference", "host": thread_detail,
                "sample": "%s %dx%d at %d spp (1/%d of the workload's spp; cost is linear in spp), reference binary "
                          "--ncores %d, render phase %.1f s (wall %.1f s)" % (
                              base, xres, yres, CPU_SAMPLE_SPP, max(1, WORKLOADS[args.workload][3] // CPU_SAMPLE_SPP),
                              threads, sec, wall)}
        if args.film_out:
            np.save(args.film_out, film.cpu().numpy() if world == 1 else host_film.numpy())
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
