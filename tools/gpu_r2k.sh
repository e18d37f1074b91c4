#!/bin/bash
# round 2, GPU call K: thread-per-read replay of the extension plans; device-phase concurrency sweep
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_sam.py tests/test_gpu_kernels.py -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]" | tail -6 > $OUT/r2k_pytest.log; cat $OUT/r2k_pytest.log
B="python bench.py --workload c3 --steps 16 --warmup 3 --no-cpu-baseline"
timeout 900 $B > $OUT/r2k_bench_c3.json 2> $OUT/r2k_bench_c3.err
EMAB_REPLAY_LANES=0 timeout 600 $B > $OUT/r2k_bench_c3_warpreplay.json 2>> $OUT/r2k_bench_c3.err
EMAB_GATE_CAPS=3,5,3 timeout 600 $B > $OUT/r2k_bench_c3_caps353.json 2>> $OUT/r2k_bench_c3.err
EMAB_GATE_CAPS=3,8,3 timeout 600 $B --workers 12 > $OUT/r2k_bench_c3_caps383.json 2>> $OUT/r2k_bench_c3.err
EMAB_GATE_CAPS=2,2,2 timeout 600 $B > $OUT/r2k_bench_c3_caps222.json 2>> $OUT/r2k_bench_c3.err
python - <<'PY'
import json
for t in ("", "_warpreplay", "_caps353", "_caps383", "_caps222"):
    try:
        d = json.load(open(f"gpurun_out/r2k_bench_c3{t}.json"))
        print(t or "default", round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items()})
    except Exception as e:
        print(t, "failed", e)
PY
