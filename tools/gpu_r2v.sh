#!/bin/bash
# round 2, GPU call V: how much of the device the seeding kernel should take when several buckets are in flight
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2v}
B="python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 600 $B > $OUT/${TAG}_bench_c3_$tag.json 2>> $OUT/${TAG}_bench_c3.err; }
run b6c3 EMAB_SEED_BPS=6
run b4c3 EMAB_SEED_BPS=4
run b3c3 EMAB_SEED_BPS=3
run b6c5 EMAB_SEED_BPS=6 EMAB_GATE_CAPS=3,5,3
run b4c5 EMAB_SEED_BPS=4 EMAB_GATE_CAPS=3,5,3
grep -i "error\|Traceback" $OUT/${TAG}_bench_c3.err | tail -8
python - <<PY
import json
for t in "b6c3 b4c3 b3c3 b6c5 b4c5".split():
    try:
        d = json.load(open(f"gpurun_out/${TAG}_bench_c3_{t}.json"))
        print(t, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k: round(v, 3) for k, v in d["device_ms_per_step"].items() if k=="ms_seed"}, "roof", round(d["roofline"]["frac"],3))
    except Exception as e:
        print(t, "failed", e)
PY
