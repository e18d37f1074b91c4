#!/bin/bash
# On the GPU box: gpu tests, then the c2 bench at several sizes of the persistent seeding grid.
OUT=gpurun_out; mkdir -p $OUT
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) | tee $OUT/sweep_pytest.log
for bps in ${@:-1 2 3 4 6}; do
  EMAB_SEED_BPS=$bps timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null > $OUT/sweep_bps$bps.json
  python - <<PY
import json
d=json.load(open("$OUT/sweep_bps$bps.json"))
print("bps=$bps", {k:round(v,3) for k,v in d["device_ms_per_step"].items()}, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3))
PY
done
