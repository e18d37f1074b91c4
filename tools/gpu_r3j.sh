#!/bin/bash
# round 2, GPU call 3j: host cloud building with compact sort keys and integer name ids: SAM parity, host profile, pinned-core e2e
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r3j}
timeout 1500 python -m pytest tests/test_gpu_sam.py tests/test_gpu_platforms.py -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]\|^\[bwa\|^\[bwt\|^\[main\]" | tail -4
B="python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline"
timeout 600 $B > $OUT/${TAG}_bench_c3_free.json 2>> $OUT/${TAG}_bench_c3.err
for th in 8 4; do timeout 600 taskset -c 0-$((th-1)) $B --threads $th > $OUT/${TAG}_bench_c3_t$th.json 2>> $OUT/${TAG}_bench_c3.err; done
EMAB_HOST_PROFILE=1 timeout 600 $B > /dev/null 2> $OUT/${TAG}_host_profile.log; grep "host profile" $OUT/${TAG}_host_profile.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_c3_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("_c3_")[1][:-5], round(d["value"]), round(d["e2e"]["value"]), [round(x, 2) for x in d["e2e"]["ms_per_step_repeats"]], {k: round(v,1) for k,v in d["e2e"]["stage_ms_per_step_summed_over_workers"].items()})
    except Exception as e:
        print(f, "failed", e)
PY
