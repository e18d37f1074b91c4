#!/bin/bash
# round 2, GPU call L: seeding grid size vs end-to-end throughput (smaller persistent grids leave room for another bucket's kernels)
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --workload c3 --steps 16 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 600 $B > $OUT/r2l_bench_c3_$tag.json 2>> $OUT/r2l_bench_c3.err; }
run m1b3 EMAB_SEED_MODE=1 EMAB_SEED_BPS=3
run m1b2 EMAB_SEED_MODE=1 EMAB_SEED_BPS=2
run m1b4 EMAB_SEED_MODE=1 EMAB_SEED_BPS=4
run m3b3 EMAB_SEED_MODE=3 EMAB_SEED_BPS=3
run m3b4 EMAB_SEED_MODE=3 EMAB_SEED_BPS=4
run m1b3c5 EMAB_SEED_MODE=1 EMAB_SEED_BPS=3 EMAB_GATE_CAPS=3,5,3
python - <<'PY'
import json
for t in ("m1b3", "m1b2", "m1b4", "m3b3", "m3b4", "m1b3c5"):
    try:
        d = json.load(open(f"gpurun_out/r2l_bench_c3_{t}.json"))
        print(t, round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items() if k in ("ms_seed",)})
    except Exception as e:
        print(t, "failed", e)
PY
