#!/bin/bash
# round 2, GPU call M: one-warp blocks for the seeding kernel: parity, then value/e2e against 128-thread blocks
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]" | tail -4 > $OUT/r2m_pytest.log; cat $OUT/r2m_pytest.log
B="python bench.py --workload c3 --steps 16 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 600 $B > $OUT/r2m_bench_c3_$tag.json 2>> $OUT/r2m_bench_c3.err; }
run w32 EMAB_SEED_BLOCK=32
run w128 EMAB_SEED_BLOCK=128
run w32c5 EMAB_SEED_BLOCK=32 EMAB_GATE_CAPS=3,5,3
run w32b4 EMAB_SEED_BLOCK=32 EMAB_SEED_BPS=4
python - <<'PY'
import json
for t in ("w32", "w128", "w32c5", "w32b4"):
    try:
        d = json.load(open(f"gpurun_out/r2m_bench_c3_{t}.json"))
        print(t, round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items() if k in ("ms_seed",)})
    except Exception as e:
        print(t, "failed", e)
PY
