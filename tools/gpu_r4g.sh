#!/bin/bash
# one GPU, the process pinned to 4 cores (a rank of a 32-core 8-GPU node): threads per bucket phase, wait mode
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r4g}
B="python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline --threads 4"
run() { tag=$1; shift; env OMP_WAIT_POLICY=passive "$@" timeout 600 taskset -c 0-3 $B > $OUT/${TAG}_bench_c3_$tag.json 2>> $OUT/${TAG}_bench_c3.err; }
run default A=1
run per1 EMAB_PHASE_THREADS=1
run per1_block EMAB_PHASE_THREADS=1 EMAB_SYNC=block
run per1_c222 EMAB_PHASE_THREADS=1 EMAB_GATE_CAPS=2,2,2
run per4_c131 EMAB_PHASE_THREADS=4 EMAB_GATE_CAPS=1,3,1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_c3_*.json")):
    d = json.load(open(f)); print(f.split("_c3_")[1][:-5], round(d["e2e"]["value"]), [round(x, 2) for x in d["e2e"]["ms_per_step_repeats"]])
PY
