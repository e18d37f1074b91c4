#!/usr/bin/env python
"""Source-line profile of one kernel: joins the per-SASS-instruction counters of an .ncu-rep (executed warp
instructions, stall samples) with the line table of the kernel's cubin (nvdisasm -g) and sums them per source line.
usage: ncu_lines.py rep kernel-regex cubin [top-n]     (cubin: cuobjdump -xelf all libema_b200.so)"""
import csv, io, re, subprocess, sys
from collections import defaultdict
rep, kre, cubin = sys.argv[1:4]
kre_ncu = kre   # "ncu-regex::cubin-regex" when the demangled (ncu) and mangled (nvdisasm) names need different patterns (templates)
if "::" in kre: kre_ncu, kre = kre.split("::", 1)
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines = []   # (file:line) per SASS instruction of the kernel, in order
cur, infn, loc = None, False, "?"
for l in dis:
    m = re.match(r"\s*\.text\.(\S+):", l) or re.match(r"\s*//-+ \.text\.(\S+)", l)
    if m:
        infn = re.search(kre, m.group(1)) is not None
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        loc = m.group(1).split("/")[-1] + ":" + m.group(2)
        if "inlined at" in l:
            pass
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(loc)
src = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre_ncu, "--launch-count", "1"],
                                                  capture_output=True, text=True).stdout)))
h = [i for i, r in enumerate(src) if "Source" in r][0]
sx = {n: i for i, n in enumerate(src[h])}
rows = [r for r in src[h + 1:] if len(r) >= len(src[h]) and r[sx["Instructions Executed"]].isdigit()]
if len(rows) == 2 * len(lines): rows = rows[:len(lines)]  # the source page repeats the listing
print(f"{len(rows)} SASS rows in the report, {len(lines)} in the cubin")
agg = defaultdict(lambda: [0, 0, 0])
for k, r in enumerate(rows):
    loc = lines[k] if k < len(lines) else "?"
    a = agg[loc]
    a[0] += int(r[sx["Instructions Executed"]]); a[1] += int(r[sx["# Samples"]] or 0); a[2] += int(r[sx["Thread Instructions Executed"]])
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{loc:28s} inst {100 * a[0] / ti:5.1f}%  samples {100 * a[1] / max(ts, 1):5.1f}%  lanes {a[2] / max(a[0], 1):4.1f}")
