#!/bin/bash
# round 2, GPU call Y: device timeline of the end-to-end pass (which stages of which buckets overlap)
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2y}
rm -f /tmp/tl.txt
EMAB_TIMELINE=/tmp/tl.txt timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline --e2e-repeats 1 > $OUT/${TAG}_bench_c3.json 2> $OUT/${TAG}_bench_c3.err
python tools/timeline.py /tmp/tl.txt 20 | tee $OUT/${TAG}_timeline.txt
cp /tmp/tl.txt $OUT/${TAG}_timeline_raw.txt
rm -f /tmp/tl.txt
EMAB_GATE_CAPS=3,5,3 EMAB_TIMELINE=/tmp/tl.txt timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline --e2e-repeats 1 > $OUT/${TAG}_bench_c3_c5.json 2>> $OUT/${TAG}_bench_c3.err
python tools/timeline.py /tmp/tl.txt 20 | tee $OUT/${TAG}_timeline_c5.txt
python - <<PY
import json
for t in ("", "_c5"):
    d = json.load(open(f"gpurun_out/${TAG}_bench_c3{t}.json"))
    print(t, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), d["e2e"]["ms_per_step_repeats"])
PY
