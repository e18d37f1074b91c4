#!/bin/bash
# round 2, GPU call 3h: one GPU, the process PINNED to the cores a rank has on a 4-GPU / 8-GPU node of 32 cores: wait modes
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r3h}
B="python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline"
run() { tag=$1; th=$2; shift; shift; env "$@" timeout 600 taskset -c 0-$((th-1)) $B --threads $th > $OUT/${TAG}_bench_c3_$tag.json 2>> $OUT/${TAG}_bench_c3.err; }
timeout 600 $B > $OUT/${TAG}_bench_c3_free.json 2>> $OUT/${TAG}_bench_c3.err
for th in 8 4; do
run t${th}_default $th A=1
run t${th}_spin $th EMAB_SYNC=spin
run t${th}_nap $th EMAB_SYNC=nap
run t${th}_block $th EMAB_SYNC=block
done
EMAB_SEED_PROF=1 timeout 600 python bench.py --workload c3 --steps 1 --warmup 1 --single-only --no-cpu-baseline 2>&1 >/dev/null | grep "seed profile" | tail -2
EMAB_HOST_PROFILE=1 timeout 600 $B > /dev/null 2> $OUT/${TAG}_host_profile.log; grep "host profile" $OUT/${TAG}_host_profile.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_c3_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("_c3_")[1][:-5], round(d["value"]), round(d["e2e"]["value"]), [round(x, 2) for x in d["e2e"]["ms_per_step_repeats"]], {k: round(v,1) for k,v in d["e2e"]["stage_ms_per_step_summed_over_workers"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
