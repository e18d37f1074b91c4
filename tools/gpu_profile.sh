#!/bin/bash
# Runs on the GPU box (under gpurun): the default bench line, the reference arm, the ncu launch list of
# the same command and one `ncu --set full` capture of the pipeline kernels.  Outputs -> gpurun_out/.
# usage: tools/gpu_profile.sh <tag> [workload]
set -u
TAG=${1:-r1}
WL=${2:-c2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 500 > $OUT/${TAG}_clocks.csv &
SMI=$!
timeout 900 python bench.py --workload $WL > $OUT/${TAG}_bench_${WL}.json 2> $OUT/${TAG}_bench_${WL}.err
tail -3 $OUT/${TAG}_bench_${WL}.err; cat $OUT/${TAG}_bench_${WL}.json
timeout 600 python bench.py --workload $WL --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_${WL}_reference.json 2>> $OUT/${TAG}_bench_${WL}.err
cat $OUT/${TAG}_bench_${WL}_reference.json
kill $SMI
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_${WL}.csv \
    python bench.py --workload $WL --steps 2 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
tail -2 $OUT/${TAG}_ncu_bench.log
# full capture of the pipeline kernels of one warm bucket: buckets one at a time (--single-only), two warm-up buckets
# = 18 launches of these kernels skipped, then the 9 of the next bucket
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_seed|k_chain|k_align1|k_rescue|k_finalize|k_em' -s 18 -c 9 \
    -f -o $OUT/${TAG}_prof_${WL} python bench.py --workload $WL --steps 1 --warmup 2 --single-only --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-300
ls -la $OUT
