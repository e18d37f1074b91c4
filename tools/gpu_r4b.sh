#!/bin/bash
# where to cut between the thread-per-task and the warp-per-task global-alignment kernels
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r4b}
B="python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu-baseline --single-only"
for w in 32 28 24 20 16; do EMAB_GLOB_WIDE_COLS=$w timeout 600 $B > $OUT/${TAG}_bench_c3_w$w.json 2>> $OUT/${TAG}_bench_c3.err; done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_c3_*.json")):
    d = json.load(open(f)); print(f.split("_c3_")[1][:-5], round(d["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items() if k in ("ms_finalize", "ms_glob_wave")})
PY
