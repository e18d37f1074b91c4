#!/bin/bash
# round 2, GPU call C: the c3 (3.1 Gbp) bench with wave timings, k_seed variants, ncu launch list + full capture on c3
OUT=gpurun_out; mkdir -p $OUT
B="python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline"
timeout 900 $B > $OUT/r2c_bench_c3.json 2> $OUT/r2c_bench_c3.err; tail -4 $OUT/r2c_bench_c3.err
EMAB_SEED_LOAD_BOTH=1 timeout 600 $B > $OUT/r2c_bench_c3_loadboth.json 2>> $OUT/r2c_bench_c3.err
EMAB_SEED_BPS=5 timeout 600 $B > $OUT/r2c_bench_c3_bps5.json 2>> $OUT/r2c_bench_c3.err
EMAB_EXT_PLAN=0 EMAB_GLOB_PLAN=0 timeout 600 $B > $OUT/r2c_bench_c3_noplan.json 2>> $OUT/r2c_bench_c3.err
python - <<'PY'
import json
for t in ("", "_loadboth", "_bps5", "_noplan"):
    try:
        d = json.load(open(f"gpurun_out/r2c_bench_c3{t}.json"))
        print(t or "default", round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items()}, d["roofline"]["frac"])
        if not t: print(d["sw"])
    except Exception as e:
        print(t, "failed", e)
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/r2c_launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --single-only > $OUT/r2c_ncu_bench.log 2>&1
tail -2 $OUT/r2c_ncu_bench.log | cut -c1-200
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_seed|k_chain|k_align1|k_rescue|k_finalize|k_ext_wave|k_glob_wave|k_ext_plan|k_glob_plan' -s 28 -c 14 \
    -f -o $OUT/r2c_prof_c3 python bench.py --workload c3 --steps 1 --warmup 2 --single-only --no-cpu-baseline > $OUT/r2c_ncu_full.log 2>&1
tail -2 $OUT/r2c_ncu_full.log | cut -c1-300
ls -la $OUT | tail -12
