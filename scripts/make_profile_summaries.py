#!/usr/bin/env python
"""Turn the ncu exports a GPU visit left in gpurun_out/ (scripts/gpu_round.sh) into the text summaries
committed under profiles/.  usage: make_profile_summaries.py <round tag, e.g. r01>"""
import csv, collections, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = "gpurun_out", "profiles"
RAW = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
       'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
       'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
       'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
       'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__t_bytes.sum',
       'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
       'smsp__warp_issue_stalled_wait_per_warp_active.pct', 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct',
       'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct',
       'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_not_selected_per_warp_active.pct',
       'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct']
for k in ("k_trace", "k_logic", "k_eval", "k_path", "k_tables", "k_vertex", "k_wshade", "k_weval", "k_wtrace"):
    raw = os.path.join(G, "full_%s_raw.csv" % k)
    if not os.path.exists(raw):
        continue
    rows = list(csv.reader(open(raw)))
    H, U, V = rows[0], rows[1], rows[2]
    out = ["ncu --set full --clock-control none --import-source on -k regex:%s -c 1 (one mid-render launch), command:" % k,
           "    python bench.py --spp 256 --steps 1 --warmup 1 --no-cpu-baseline      (BASELINE configs[1] scene, 1024x1024; round 1 used --spp 128)",
           "kernel: " + V[H.index("Kernel Name")], ""]
    for m in RAW:
        if m in H:
            out.append("%-78s %s %s" % (m, V[H.index(m)], U[H.index(m)]))
    out += ["", "per CUDA source line (samples, instructions, active threads per warp instruction):"]
    src = os.path.join(G, "full_%s_src.csv" % k)
    if os.path.exists(src):
        out.append(subprocess.run([sys.executable, "scripts/ncu_lines.py", src, "25"], stdout=subprocess.PIPE, text=True).stdout)
    open(os.path.join(P, "%s_%s_ncu_full.txt" % (tag, k)), "w").write("\n".join(out))
# launch list: per-kernel totals and shares
ll = os.path.join(G, "launches.csv")
if os.path.exists(ll):
    rows = list(csv.reader(open(ll)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hi]
    ix = {h: i for i, h in enumerate(H)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(H) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
        if name.startswith("cub::"):
            name = "cub::DeviceRadixSort*"
        if name.startswith("at::"):
            name = "torch fill (L2 flush / film clear)"
        agg.setdefault(name, []).append(float(r[ix["Metric Value"]].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = ["ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 on `python bench.py --steps 1 --warmup 1 --no-cpu-baseline`",
           "(first 4000 launches: the counting pass -- STATS instantiations <.., 1, ..> -- and the start of the next pass;",
           " per-launch times are serialised and cold-cache: only the kernels' SHARES are meaningful)", "",
           "%-34s %7s %12s %10s %7s" % ("kernel", "launches", "total ms", "avg us", "share")]
    for k, v in agg.items():
        out.append("%-34s %7d %12.2f %10.1f %6.1f%%" % (k[:34], len(v), sum(v) / 1e6, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    open(os.path.join(P, "%s_launch_shares.txt" % tag), "w").write("\n".join(out) + "\n")
    import shutil
    shutil.copyfile(ll, os.path.join(P, "%s_launches.csv" % tag))
print(sorted(os.listdir(P)))
