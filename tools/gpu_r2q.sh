#!/bin/bash
# round 2, GPU call Q: ncu --set full of the seeding kernel (default form) on the 3.1 Gbp index, third bucket
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2q}
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_seed_rq|k_seed_hot|k_seed_wide' -s 2 -c 1 \
    -f -o $OUT/${TAG}_prof_seed python bench.py --workload c3 --steps 1 --warmup 2 --single-only --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
tail -2 $OUT/${TAG}_ncu_full.log | cut -c1-300
ls -la $OUT/${TAG}_prof_seed.ncu-rep
