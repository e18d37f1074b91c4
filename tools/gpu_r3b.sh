#!/bin/bash
# round 2, GPU call 3b: replay without DP rows in shared memory; glob wave with a ring row; ext waves in two length classes side by side
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r3b}
timeout 1500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_kernels.py tests/test_gpu_edge.py -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]\|^\[bwa\|^\[bwt\|^\[main\]" | tail -5
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_c3.json"))
print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d["e2e"]["ms_per_step_repeats"], {k: round(v, 3) for k, v in d["device_ms_per_step"].items()}, d["sw"]["extend_calls_inline_per_step"])
PY
