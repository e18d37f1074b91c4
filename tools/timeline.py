#!/usr/bin/env python
"""Reads an EMAB_TIMELINE file (one line per device-pipeline call: ctx, then twelve stage-event times in ms on the device's
clock: seed0 seed1 chain0 chain1 ext0 ext1 align1_end rescue_end fin0 glob0 glob1 fin_end) and prints, for the last N calls
(the timed end-to-end pass), how the buckets' stages overlap.  usage: timeline.py file N"""
import sys
import numpy as np
rows = [l.split() for l in open(sys.argv[1]) if l.strip()]
N = int(sys.argv[2])
rows = rows[-N:]
T = np.array([[float(x) for x in r[1:]] for r in rows])
ctx = [r[0] for r in rows]
t0 = T[:, 0].min()
T = np.where(T > 0, T - t0, 0)
names = ["seed", "chain", "ext_waves", "align1(rest)", "rescue", "gap", "finalize"]
span = T[:, 11].max() - T[:, 0].min()
print(f"{N} buckets, span {span:.2f} ms = {span / N:.2f} ms per bucket; contexts {len(set(ctx))}")
seg = {"seed": (0, 1), "gap seed->chain": (1, 2), "chain": (2, 3), "align1": (3, 6), "rescue": (6, 7), "gap rescue->finalize": (7, 8), "finalize": (8, 11)}
for k, (a, b) in seg.items():
    d = T[:, b] - T[:, a]
    print(f"  {k:22s} mean {d.mean():7.3f}  min {d.min():7.3f}  max {d.max():7.3f}")
tot = T[:, 11] - T[:, 0]
print(f"  bucket first-to-last   mean {tot.mean():7.3f}  min {tot.min():7.3f}  max {tot.max():7.3f}")
# how many buckets are inside [seed0, fin_end] at a time, and how many seeds at a time
ev = sorted([(t, 1) for t in T[:, 0]] + [(t, -1) for t in T[:, 11]])
cur = 0; last = ev[0][0]; acc = {}
for t, d in ev:
    acc[cur] = acc.get(cur, 0) + (t - last); last = t; cur += d
print("  time with k buckets in the device phase:", {k: round(v, 2) for k, v in sorted(acc.items())})
ev = sorted([(t, 1) for t in T[:, 0]] + [(t, -1) for t in T[:, 1]])
cur = 0; last = ev[0][0]; acc = {}
for t, d in ev:
    acc[cur] = acc.get(cur, 0) + (t - last); last = t; cur += d
print("  time with k seeding kernels running:", {k: round(v, 2) for k, v in sorted(acc.items())})
order = np.argsort(T[:, 0])
print("  first 10 buckets (ms): seed0 seed1 | chain0 chain1 | ext0 ext1 align1_end | rescue_end | fin0 glob0 glob1 fin_end")
for i in order[:10]:
    print("   ", ctx[i][-5:], " ".join(f"{v:7.2f}" for v in T[i]))
