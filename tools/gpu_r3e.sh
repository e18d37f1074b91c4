#!/bin/bash
# round 2, GPU call 3e: buckets in the device phase x seeding footprint, median of three end-to-end repetitions each
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r3e}
B="python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 600 $B > $OUT/${TAG}_bench_c3_$tag.json 2>> $OUT/${TAG}_bench_c3.err; }
for caps in 2 3 4; do for bps in 6 5 4; do run c${caps}b${bps} EMAB_SEED_BPS=$bps EMAB_GATE_CAPS=3,$caps,3; done; done
run c3b6w12 EMAB_SEED_BPS=6 EMAB_BENCH_WORKERS=12
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_c3_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("_c3_")[1][:-5], round(d["value"]), round(d["e2e"]["value"]), [round(x, 2) for x in d["e2e"]["ms_per_step_repeats"]], round(d["device_ms_per_step"]["ms_seed"], 2))
    except Exception as e:
        print(f, "failed", e)
PY
