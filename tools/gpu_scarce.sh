#!/bin/bash
# On a 1-GPU box: emulate the 8-cores-per-GPU host of the multi-GPU boxes (taskset to 8 cores, 8 host threads) and
# try the knobs that matter when cores are scarce.  usage: tools/gpu_scarce.sh
OUT=gpurun_out; mkdir -p $OUT
run() {  # name, workers, env...
  name=$1; w=$2; shift 2
  env "$@" taskset -c 0-7 timeout 600 python bench.py --steps 25 --warmup 3 --workers $w --threads 8 --no-cpu-baseline 2> $OUT/scarce.err > $OUT/scarce.json
  python - <<PY
import json
d=[json.loads(l) for l in open("$OUT/scarce.json") if l.startswith("{")][-1]
print("$name workers=$w", "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2), "value", round(d["value"]))
PY
}
run default 8 A=1
run block 8 EMAB_SYNC=block
run block+passive 8 EMAB_SYNC=block OMP_WAIT_POLICY=passive
run block+passive+caps232 8 EMAB_SYNC=block OMP_WAIT_POLICY=passive EMAB_GATE_CAPS=2,3,2
run block+passive+caps222+w6 6 EMAB_SYNC=block OMP_WAIT_POLICY=passive EMAB_GATE_CAPS=2,2,2
run block+caps232 8 EMAB_SYNC=block EMAB_GATE_CAPS=2,3,2
run passive 8 OMP_WAIT_POLICY=passive
