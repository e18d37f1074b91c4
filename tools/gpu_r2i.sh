#!/bin/bash
# round 2, GPU call I: device SAM formatter + text upload path: the whole gpu suite, then c3 bench (wait modes), host profile
OUT=gpurun_out; mkdir -p $OUT
timeout 3000 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]" | tail -25 > $OUT/r2i_pytest.log; cat $OUT/r2i_pytest.log
B="python bench.py --workload c3 --steps 12 --warmup 3 --no-cpu-baseline"
EMAB_HOST_PROFILE=1 timeout 900 $B > $OUT/r2i_bench_c3.json 2> $OUT/r2i_bench_c3.err; grep "host profile" $OUT/r2i_bench_c3.err | tail -2
EMAB_SYNC=spin timeout 600 $B > $OUT/r2i_bench_c3_spin.json 2>> $OUT/r2i_bench_c3.err
EMAB_SYNC=block timeout 600 $B > $OUT/r2i_bench_c3_block.json 2>> $OUT/r2i_bench_c3.err
python - <<'PY'
import json
for t in ("", "_spin", "_block"):
    try:
        d = json.load(open(f"gpurun_out/r2i_bench_c3{t}.json"))
        print(t or "default", round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items()})
        if not t: print(d["host_ms_per_step"], d["e2e"])
    except Exception as e:
        print(t, "failed", e)
PY
