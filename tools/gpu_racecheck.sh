#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r4k}
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/${TAG}_racecheck_seed.log \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_edge.py -x -q -m gpu -k "smem_default or long_interval" > $OUT/${TAG}_racecheck_seed.out 2>&1
echo "rc $?"; tail -2 $OUT/${TAG}_racecheck_seed.out; grep -c "hazard" $OUT/${TAG}_racecheck_seed.log; tail -4 $OUT/${TAG}_racecheck_seed.log
timeout 300 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/${TAG}_racecheck_pipe.log \
    python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "golden" > $OUT/${TAG}_racecheck_pipe.out 2>&1
echo "rc $?"; tail -2 $OUT/${TAG}_racecheck_pipe.out; tail -3 $OUT/${TAG}_racecheck_pipe.log
