#!/usr/bin/env python
"""Digest of an .ncu-rep: key raw metrics per kernel + the hottest SASS segments (runs of instructions
with the same execution count) with lanes-per-instruction and stall samples.  usage: ncu_digest.py rep [kernel-regex]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else None
WANT = ["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","lts__t_sectors.sum","lts__t_sector_hit_rate.pct",
 "sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__grid_size","launch__block_size",
 "smsp__inst_executed.sum","smsp__thread_inst_executed_per_inst_executed.ratio","smsp__issue_active.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active","sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
 "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed","l1tex__t_sector_hit_rate.pct",
 "smsp__warps_eligible.avg.per_cycle_active","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
 "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio","smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio","smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio","smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
 "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio"]
def run(args):
    return subprocess.run(["ncu", "-i", rep] + args, capture_output=True, text=True).stdout
raw = list(csv.reader(io.StringIO(run(["--page", "raw", "--csv"]))))
hdr, units, data = raw[0], raw[1], raw[2:]
ix = {h: i for i, h in enumerate(hdr)}
names = [r[ix["Kernel Name"]].split("(")[0] for r in data]
print("| metric | " + " | ".join(names) + " |"); print("|---|" + "---|" * len(names))
for w in WANT:
    if w in ix: print(f"| {w} [{units[ix[w]]}] | " + " | ".join(r[ix[w]] for r in data) + " |")
if kre:
    src = list(csv.reader(io.StringIO(run(["--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-count", "1"]))))
    h = [i for i, r in enumerate(src) if "Source" in r][0]
    sh = src[h]; sx = {n: i for i, n in enumerate(sh)}
    rows = [r for r in src[h + 1:] if len(r) >= len(sh) and r[sx["Instructions Executed"]].isdigit()]
    tot = sum(int(r[sx["Instructions Executed"]]) for r in rows); samp = sum(int(r[sx["# Samples"]] or 0) for r in rows)
    print(f"\nSASS: {len(rows)} instructions, {tot} warp-instructions executed, {samp} samples")
    seg = []; prev = None
    for k, r in enumerate(rows):
        ie = int(r[sx["Instructions Executed"]]); te = int(r[sx["Thread Instructions Executed"]]); sp = int(r[sx["# Samples"]] or 0)
        if prev is None or ie != prev: seg.append([ie, te, k, k, [r[sx["Source"]].strip()], sp]); prev = ie
        else: seg[-1][1] += te; seg[-1][3] = k; seg[-1][4].append(r[sx["Source"]].strip()); seg[-1][5] += sp
    for s in sorted(seg, key=lambda s: -s[5])[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]:
        n = s[3] - s[2] + 1
        ops = [(x.split()[0] if not x.startswith("@") else x.split()[1]).split(".")[0] for x in s[4]]
        print(f"sass[{s[2]:4d}-{s[3]:4d}] n={n:3d} exec={s[0]:>9} inst={100*s[0]*n/tot:5.1f}% samples={100*s[5]/max(samp,1):5.1f}% lanes={s[1]/max(1,s[0]*n):5.1f} {Counter(ops).most_common(5)}")
