#!/bin/bash
# round 2, GPU call E: quad seeding kernel: parity tests, then c3 bench for quad (8 and 6 blocks/SM) and thread mode
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py tests/test_gpu_sam.py -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa" | tail -8 > $OUT/r2e_pytest.log; cat $OUT/r2e_pytest.log
B="python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline"
timeout 900 $B > $OUT/r2e_bench_c3_quad8.json 2> $OUT/r2e_bench_c3.err; tail -2 $OUT/r2e_bench_c3.err
EMAB_SEED_BPS=6 timeout 600 $B > $OUT/r2e_bench_c3_quad6.json 2>> $OUT/r2e_bench_c3.err
EMAB_SEED_BPS=10 timeout 600 $B > $OUT/r2e_bench_c3_quad10.json 2>> $OUT/r2e_bench_c3.err
EMAB_SEED_MODE=1 timeout 600 $B > $OUT/r2e_bench_c3_thread.json 2>> $OUT/r2e_bench_c3.err
python - <<'PY'
import json
for t in ("quad8", "quad6", "quad10", "thread"):
    try:
        d = json.load(open(f"gpurun_out/r2e_bench_c3_{t}.json"))
        print(t, round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items()}, round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(t, "failed", e)
PY
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_seed_quad|k_align1|k_glob_wave|k_ext_wave' -s 10 -c 5 \
    -f -o $OUT/r2e_prof_c3 python bench.py --workload c3 --steps 1 --warmup 2 --single-only --no-cpu-baseline > $OUT/r2e_ncu_full.log 2>&1
tail -2 $OUT/r2e_ncu_full.log | cut -c1-200
