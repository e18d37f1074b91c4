#!/bin/bash
# round 2, GPU call H: entry-parallel quad seeding (mode 3): parity tests, c3 bench sweep, target-regime test
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_sam.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]" | tail -8 > $OUT/r2h_pytest.log; cat $OUT/r2h_pytest.log
B="python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline"
for bps in 5 6 4 3; do EMAB_SEED_BPS=$bps timeout 600 $B > $OUT/r2h_bench_c3_wide$bps.json 2>> $OUT/r2h_bench_c3.err; done
EMAB_SEED_MODE=1 timeout 600 $B > $OUT/r2h_bench_c3_thread.json 2>> $OUT/r2h_bench_c3.err
python - <<'PY'
import json
for t in ("wide5", "wide6", "wide4", "wide3", "thread"):
    try:
        d = json.load(open(f"gpurun_out/r2h_bench_c3_{t}.json"))
        print(t, round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items()}, round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(t, "failed", e)
PY
