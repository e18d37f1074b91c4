#!/bin/bash
# On the GPU box: everything the round's evidence needs from one call — gpu tests, smoke, the SW
# microbenchmark (+ ncu of the thread-per-task kernel), the c2 bench with its reference arm, the ncu
# launch list and a full capture of the pipeline kernels.  usage: tools/gpu_r1e.sh <tag>
TAG=${1:-r1e}
OUT=gpurun_out
mkdir -p $OUT
(nproc; lscpu | head -20; nvidia-smi -L) > $OUT/${TAG}_box.txt 2>&1
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) | tee $OUT/${TAG}_pytest_gpu.log
(timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3) | tee $OUT/${TAG}_smoke.log
timeout 900 python bench_sw.py --mode 0 > $OUT/${TAG}_bench_sw_lanes.jsonl 2> $OUT/${TAG}_bench_sw.err; cut -c1-700 $OUT/${TAG}_bench_sw_lanes.jsonl; tail -5 $OUT/${TAG}_bench_sw.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extend_lanes -s 2 -c 1 -f -o $OUT/${TAG}_prof_sw_lanes \
    python bench_sw.py --mode 0 --lengths 151 --no-cpu --reps 1 --warmup 1 > $OUT/${TAG}_ncu_sw.log 2>&1
tail -3 $OUT/${TAG}_ncu_sw.log
bash tools/gpu_profile.sh $TAG c2
