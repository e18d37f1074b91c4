#!/bin/bash
# round 2, GPU call 3a: cooperative handling of unplanned extensions in the replay: parity + bench; ncu of the non-seeding kernels
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r3a}
timeout 1500 python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_c3.json"))
print(round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d["e2e"]["ms_per_step_repeats"], {k: round(v, 3) for k, v in d["device_ms_per_step"].items()}, d["sw"]["extend_calls_inline_per_step"])
PY
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_align1_replay|k_glob_wave|k_chain|k_ext_wave|k_rescue_sw|k_finalize' -s 12 -c 7 \
    -f -o $OUT/${TAG}_prof_sw python bench.py --workload c3 --steps 1 --warmup 2 --single-only --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
tail -1 $OUT/${TAG}_ncu_full.log | cut -c1-200
