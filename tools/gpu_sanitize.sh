#!/bin/bash
# compute-sanitizer memcheck over the seeding kernel tests and one pipeline test (the kernels added this round)
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r4j}
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/${TAG}_memcheck_seed.log \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_edge.py -x -q -m gpu -k "smem or long_interval" > $OUT/${TAG}_memcheck_seed.out 2>&1
echo "rc $?"; tail -3 $OUT/${TAG}_memcheck_seed.out; grep -c "Invalid\|out of bounds\|misaligned" $OUT/${TAG}_memcheck_seed.log; tail -3 $OUT/${TAG}_memcheck_seed.log
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/${TAG}_memcheck_pipe.log \
    python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "golden" > $OUT/${TAG}_memcheck_pipe.out 2>&1
echo "rc $?"; tail -3 $OUT/${TAG}_memcheck_pipe.out; tail -3 $OUT/${TAG}_memcheck_pipe.log
