#!/bin/bash
# On the GPU box: kernel parity tests, the SW microbenchmark (thread-per-task kernel) and an ncu capture of it.
# usage: tools/gpu_sw.sh <tag> [lengths]
TAG=${1:-r1}
LENS=${2:-101,151,251}
OUT=gpurun_out
mkdir -p $OUT
(timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -15) | tee $OUT/${TAG}_pytest_kernels.log
timeout 900 python bench_sw.py --mode 0 --lengths $LENS > $OUT/${TAG}_bench_sw_lanes.jsonl 2> $OUT/${TAG}_bench_sw.err; cut -c1-600 $OUT/${TAG}_bench_sw_lanes.jsonl | grep -v int_pipe; tail -5 $OUT/${TAG}_bench_sw.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend_lanes -s 2 -c 1 -f -o $OUT/${TAG}_prof_sw_lanes \
    python bench_sw.py --mode 0 --lengths 151 --no-cpu --reps 1 --warmup 1 > $OUT/${TAG}_ncu_sw.log 2>&1
tail -3 $OUT/${TAG}_ncu_sw.log
