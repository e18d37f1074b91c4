#!/bin/bash
# round 2, GPU call B: index-build parity again, the wave-scheduled SW in the pipeline (tests + bench)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_index_build.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r2b_pytest_index.log; cat gpurun_out/r2b_pytest_index.log
timeout 1200 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_kernels.py tests/test_gpu_sam.py -x -q -m gpu -s 2>&1 | tail -25 > gpurun_out/r2b_pytest_pipeline.log; cat gpurun_out/r2b_pytest_pipeline.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2b_bench_c2.json 2> gpurun_out/r2b_bench_c2.err; tail -3 gpurun_out/r2b_bench_c2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench_c2.json'))
print({k:d[k] for k in ('value','ms_per_step','device_ms_per_step','host_ms_per_step')}, d['e2e']['value'])
PY
