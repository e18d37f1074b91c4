#!/usr/bin/env python
"""bench_sw.py — BASELINE.json configs[4]: batched banded-SW microbenchmark.

ksw_extend2 (bwa/ksw.c:416) on 101/151/251-bp queries, band w=100, tasks built as SURVEY.md §8d
config 5 prescribes (target = query with 1 % subst / 0.1 % ins / 0.1 % del, padded with random bases
to tlen = qlen+100; a=1,b=4,o=6,e=1, end_bonus=5, zdrop=100, h0=19), >= 1 M tasks per length
(131 072 distinct tasks replicated 8x so the generator stays fast; inputs are resident in HBM).

Per length one JSON line: GCUPS counting the DP cells the reference loop visits (the kernel counts
them; asserted equal to the oracle's count on a sample) and the nominal qlen x tlen figure, the
integer roofline (visited cells x 15 int ops, SURVEY.md §8d, against the integer-pipe throughput
measured on this GPU by emab_int_peak), and the CPU reference beside it (the compiled reference's
own ksw_extend2 on all host threads, on a sample).  Results of a sample are checked bit-exact
against the oracle before anything is timed.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from tools import synth  # noqa: E402

OPS_PER_CELL = 15  # bwa/ksw.c:467-483, SURVEY.md §8d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lengths", default="101,151,251")
    ap.add_argument("--distinct", type=int, default=131072)
    ap.add_argument("--replicate", type=int, default=8)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--mode", type=int, default=0, help="0 thread-per-task kernel, 1 warp-per-task kernel")
    ap.add_argument("--check", type=int, default=4096, help="tasks compared with the oracle before timing")
    ap.add_argument("--cpu-sample", type=int, default=20000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    import ema_b200
    import helpers

    ctx = ema_b200.Context(None)
    ema_b200.set_sw_mode(ctx, args.mode)
    # ---- integer-pipe peaks of this GPU
    kinds = {0: "add.s32 (IADD3)", 1: "max.s32 (ptxas fuses pairs into VIMNMX3; counted per max)", 2: "viaddmax (VIADDMNMX)",
             3: "vimax3 (VIMNMX3)", 4: "mad.lo (IMAD, FMA pipe)", 5: "add + mad (both pipes)", 6: "viaddmax.s16x2",
             7: "add + viaddmax", 8: "lop3", 9: "shf", 10: "prmt", 11: "max/min alternating (VIMNMX)", 12: "viaddmax + mad",
             13: "lop3 + viaddmax", 14: "lop3 + mad", 15: "min/max + viaddmax", 16: "vimax3 + mad", 17: "vimax3 + lop3", 18: "lop3 + shf",
             19: "prmt + viaddmax"}
    peaks = {}
    for k, name in kinds.items():
        g, ms = ema_b200.int_peak(ctx, k, 4000)
        peaks[name] = g
    inst_peak = max(v for n, v in peaks.items() if not n.startswith("max.s32"))  # lane-instructions/s, best mix
    if os.environ.get("EMAB_PEAKS_ONLY"):
        print(json.dumps(peaks, indent=1))
        return
    print(json.dumps({"int_pipe_peaks_Gops": peaks, "peak_used_Gops": inst_peak,
                      "peak_definition": "highest measured rate of simple int32 lane-instructions per second over the kinds above (one DPX instruction counted as ONE op)"}))

    P = helpers.port()
    R = helpers.ref() if helpers.have_ref() else None
    cores = os.cpu_count() or 1
    for qlen in [int(x) for x in args.lengths.split(",")]:
        q, t = synth.extend_tasks(args.distinct, qlen)
        h0 = np.full(args.distinct, 19, np.int32)
        # parity on a sample, through the host-buffer batch entry point
        nchk = min(args.check, args.distinct)
        got, cells = ema_b200.extend_batch(ctx, list(q[:nchk]), list(t[:nchk]), h0[:nchk])
        want, cells_cpu = helpers.sw_extend(P, "orc", list(q[:nchk]), list(t[:nchk]), h0[:nchk])
        assert np.array_equal(got, want) and cells == cells_cpu, f"qlen {qlen}: GPU ksw_extend2 differs from the oracle"
        # resident, replicated
        qq = np.tile(q, (args.replicate, 1))
        tt = np.tile(t, (args.replicate, 1))
        hh = np.tile(h0, args.replicate)
        n = ema_b200.extend_resident_load(ctx, qq, tt, hh)
        ema_b200.extend_resident_run(ctx, n, reps=args.warmup, want_out=False)
        out, vis, ms = ema_b200.extend_resident_run(ctx, n, reps=args.reps)
        assert np.array_equal(out[:nchk], want), "resident run differs from the oracle"
        assert np.array_equal(out[:args.distinct], out[-args.distinct:]), "replicas disagree"
        nominal = float(n) * qlen * (qlen + 100)
        gcups_vis = vis / (ms * 1e-3) / 1e9
        line = {"metric": "banded-SW GCUPS (ksw_extend2)", "qlen": qlen, "tlen": qlen + 100, "tasks": n, "w": 100, "kernel_mode": args.mode,
                "ms_per_launch": ms, "visited_cells_per_task": vis / n, "gcups_visited": gcups_vis,
                "gcups_nominal": nominal / (ms * 1e-3) / 1e9, "visited_over_nominal": vis / nominal,
                "roofline": {"bound": "int-alu", "achieved": gcups_vis * OPS_PER_CELL, "peak": inst_peak, "unit": "Gop/s",
                             "frac": gcups_vis * OPS_PER_CELL / inst_peak, "ops_per_cell": OPS_PER_CELL},
                "parity": f"{nchk} tasks bit-exact vs oracle (6 outputs + visited cells)"}
        if not args.no_cpu:
            ns = min(args.cpu_sample, args.distinct)
            lib, pre, kind = (R, "ref", "reference") if R is not None else (P, "orc", "port")
            t0 = time.time()
            helpers.sw_extend(lib, pre, list(q[:ns]), list(t[:ns]), h0[:ns], threads=cores)
            dt = time.time() - t0
            _, c_cpu = helpers.sw_extend(P, "orc", list(q[:ns]), list(t[:ns]), h0[:ns], threads=cores)
            line["cpu_baseline"] = {"value": c_cpu / dt / 1e9, "unit": "GCUPS visited", "cores": cores, "kind": kind,
                                    "sample": f"{ns} tasks, {cores} threads, compiled reference ksw_extend2" if R is not None else f"{ns} tasks, oracle port"}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
