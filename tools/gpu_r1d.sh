#!/bin/bash
# On the GPU box: full gpu test suite, smoke, the SW microbenchmark and an ncu capture of the thread-per-task
# extension kernel.  usage: tools/gpu_r1d.sh <tag>
TAG=${1:-r1d}
OUT=gpurun_out
mkdir -p $OUT
nproc > $OUT/${TAG}_nproc.txt; lscpu | head -20 >> $OUT/${TAG}_nproc.txt
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) | tee $OUT/${TAG}_pytest_gpu.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) | tee $OUT/${TAG}_smoke.log
timeout 900 python bench_sw.py --mode 0 > $OUT/${TAG}_bench_sw_lanes.jsonl 2> $OUT/${TAG}_bench_sw.err; cat $OUT/${TAG}_bench_sw_lanes.jsonl; tail -5 $OUT/${TAG}_bench_sw.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend_lanes -s 2 -c 1 -f -o $OUT/${TAG}_prof_sw_lanes \
    python bench_sw.py --mode 0 --lengths 151 --no-cpu --reps 1 --warmup 1 > $OUT/${TAG}_ncu_sw.log 2>&1
tail -3 $OUT/${TAG}_ncu_sw.log
