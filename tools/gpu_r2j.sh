#!/bin/bash
# round 2, GPU call J: device bucket parser: gpu suite, c3 bench + host profile
OUT=gpurun_out; mkdir -p $OUT
timeout 3000 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]" | tail -25 > $OUT/r2j_pytest.log; cat $OUT/r2j_pytest.log
B="python bench.py --workload c3 --steps 12 --warmup 3 --no-cpu-baseline"
EMAB_HOST_PROFILE=1 timeout 900 $B > $OUT/r2j_bench_c3.json 2> $OUT/r2j_bench_c3.err; grep "host profile" $OUT/r2j_bench_c3.err | tail -2
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2j_bench_c3.json"))
print(round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items()})
print(d["host_ms_per_step"], d["e2e"])
PY
