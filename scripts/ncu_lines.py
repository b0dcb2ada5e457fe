#!/usr/bin/env python
"""Summarise an ncu `--page source --print-source cuda,sass --csv` dump per CUDA source line:
samples, instructions executed, average active threads.  usage: ncu_lines.py dump.csv [topN]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
cur_fn = ""
FILTER = sys.argv[3] if len(sys.argv) > 3 else ""
agg = {}
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name":
        cur_fn = r[1]; continue
    if hdr is None: continue
    if FILTER and FILTER not in cur_fn: continue
    if not r[0].isdigit(): continue
    # source-line aggregate rows have '-' for Address
    if r[2] != "-": continue
    idx = {h: i for i, h in enumerate(hdr)}
    try:
        samples = int(r[idx["# Samples"]]); inst = int(r[idx["Instructions Executed"]])
        tinst = int(r[idx["Thread Instructions Executed"]])
    except ValueError:
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, 0, 0, r[1].strip()[:110]])
    a[0] += samples; a[1] += inst; a[2] += tinst
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values()); tot_t = sum(a[2] for a in agg.values())
print("total samples %d, warp-instructions %d, avg active threads %.2f" % (tot_s, tot_i, tot_t / max(1, tot_i)))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  thr %5.1f  %s:%d  %s" % (100.0 * a[0] / tot_s, 100.0 * a[1] / tot_i, a[2] / max(1, a[1]), key[0], key[1], a[3]))
