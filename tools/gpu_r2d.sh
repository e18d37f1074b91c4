#!/bin/bash
# round 2, GPU call D: glob wave v2 + ext wave classes: parity tests, then the c3 bench
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_sam.py -x -q -m gpu -s 2>&1 | grep -v "^BWA\|^Processing\|M::bwa" | tail -12 > $OUT/r2d_pytest.log; cat $OUT/r2d_pytest.log
B="python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline"
timeout 900 $B > $OUT/r2d_bench_c3.json 2> $OUT/r2d_bench_c3.err; tail -3 $OUT/r2d_bench_c3.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2d_bench_c3.json"))
print(round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items()}, d["roofline"]["frac"])
print({k: (round(v, 1) if isinstance(v, float) else v) for k, v in d["sw"].items() if k != "note"})
print(d["host_ms_per_step"], d["e2e"]["stage_ms_per_step_summed_over_workers"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $OUT/r2d_launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --single-only > $OUT/r2d_ncu_bench.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2d_launches_c3.csv", errors="replace")))
h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
ix = {n: i for i, n in enumerate(rows[h])}
agg = collections.OrderedDict()
for r in rows[h + 1:]:
    if len(r) <= ix["Metric Value"]: continue
    k = r[ix["Kernel Name"]].split("(")[0]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[ix["Metric Value"]].replace(",", ""))
for k, (n, t) in agg.items(): print(f"{k:60s} n={n:4d} total={t/1e6:9.3f} ms  avg={t/n/1e3:9.1f} us")
PY
