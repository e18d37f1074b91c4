#!/bin/bash
# round 2, GPU call A: index-build parity, then the target regime (1 Gbp, then 3.1 Gbp) end to end against the reference
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2a_box.txt; nproc >> gpurun_out/r2a_box.txt; free -g >> gpurun_out/r2a_box.txt
timeout 900 python -m pytest tests/test_index_build.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2a_pytest_index.log
cat gpurun_out/r2a_pytest_index.log
timeout 600 python tools/big_run.py --config g1 --buckets 1 > gpurun_out/r2a_big_g1.json 2> gpurun_out/r2a_big_g1.err; tail -5 gpurun_out/r2a_big_g1.err; cat gpurun_out/r2a_big_g1.json
rm -rf /tmp/emab_data/big_g1
timeout 1200 python tools/big_run.py --config c3 --buckets 2 --density > gpurun_out/r2a_big_c3.json 2> gpurun_out/r2a_big_c3.err; tail -8 gpurun_out/r2a_big_c3.err; cat gpurun_out/r2a_big_c3.json
