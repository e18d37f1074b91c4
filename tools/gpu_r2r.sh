#!/bin/bash
# round 2, GPU call R: the request-loop seeding kernel (seed_rq.cuh): parity, then c3 bench per variant
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2r}
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]" | tail -25 > $OUT/${TAG}_pytest.log; cat $OUT/${TAG}_pytest.log
B="python bench.py --workload c3 --steps 16 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 600 $B > $OUT/${TAG}_bench_c3_$tag.json 2>> $OUT/${TAG}_bench_c3.err; }
run rq6 EMAB_SEED_BPS=6
run rq5 EMAB_SEED_BPS=5
run rq7 EMAB_SEED_BPS=7
run rq4 EMAB_SEED_BPS=4
grep -i "error\|Traceback" $OUT/${TAG}_bench_c3.err | tail -8
python - <<PY
import json
for t in ("rq6", "rq5", "rq7", "rq4"):
    try:
        d = json.load(open(f"gpurun_out/${TAG}_bench_c3_{t}.json"))
        print(t, round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items()}, d["roofline"]["algorithmic_bytes_per_launch"])
    except Exception as e:
        print(t, "failed", e)
PY
