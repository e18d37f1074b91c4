#!/bin/bash
# ncu launch list (gpu__time_duration per launch, cold-cache and serialised: shares only) of the bench command on a prepared data set
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2z}
timeout 900 python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --single-only > /dev/null 2>&1   # builds the index and the buckets
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --single-only > $OUT/${TAG}_ncu_bench.log 2>&1
tail -2 $OUT/${TAG}_ncu_bench.log | cut -c1-200
wc -l $OUT/${TAG}_launches_c3.csv
