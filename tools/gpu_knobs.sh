#!/bin/bash
IFS_OLD=$IFS
# On the GPU box: the c2 bench under combinations of the tuning knobs given as "VAR=val+VAR=val" arguments.
OUT=gpurun_out; mkdir -p $OUT
i=0
for combo in "$@"; do
  i=$((i+1))
  env $(echo $combo | tr '+' ' ') timeout 600 python bench.py --steps ${STEPS:-10} --warmup 3 --workers ${WORKERS:-8} --threads ${THREADS:-0} --no-cpu-baseline 2> $OUT/knob_$i.err > $OUT/knob_$i.json; grep "host profile" $OUT/knob_$i.err
  python - <<PY
import json
d=json.load(open("$OUT/knob_$i.json"))
print("$combo", {k[3:] if k.startswith("ms_") else k:round(v,2) for k,v in d["device_ms_per_step"].items()}, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2))
PY
done
