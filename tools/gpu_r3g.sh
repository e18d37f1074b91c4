#!/bin/bash
# round 2, GPU call 3g: one GPU with the host threads a rank has on an 8-GPU / 4-GPU node (4 / 8): phase caps
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r3g}
timeout 1500 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -2
B="python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline --single-only-skip"
B="python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline"
run() { tag=$1; th=$2; shift; shift; env "$@" timeout 600 $B --threads $th > $OUT/${TAG}_bench_c3_$tag.json 2>> $OUT/${TAG}_bench_c3.err; }
run t24 24 A=1
for th in 8 4; do
run t${th}_c333 $th EMAB_GATE_CAPS=3,3,3
run t${th}_c232 $th EMAB_GATE_CAPS=2,3,2
run t${th}_c132 $th EMAB_GATE_CAPS=1,3,2
run t${th}_c122 $th EMAB_GATE_CAPS=1,2,2
run t${th}_c131 $th EMAB_GATE_CAPS=1,3,1
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_c3_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("_c3_")[1][:-5], round(d["value"]), round(d["e2e"]["value"]), [round(x, 2) for x in d["e2e"]["ms_per_step_repeats"]], {k: round(v,1) for k,v in d["e2e"]["stage_ms_per_step_summed_over_workers"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
