#!/bin/bash
# round 2, GPU call F: the whole -m gpu suite (new: edge cases, platform sweep incl. cpt and 100 Mbp, -d rewrite, streaming -1), then
# the c3 bench default + seed occupancy sweep (is k_seed bound by its longest chains?)
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]" | tail -15 > $OUT/r2f_pytest.log; cat $OUT/r2f_pytest.log
B="python bench.py --workload c3 --steps 6 --warmup 3 --no-cpu-baseline"
for bps in 6 4 3 2; do EMAB_SEED_BPS=$bps timeout 600 $B > $OUT/r2f_bench_c3_bps$bps.json 2>> $OUT/r2f_bench_c3.err; done
EMAB_SEED_MODE=2 EMAB_SEED_BPS=8 timeout 600 $B > $OUT/r2f_bench_c3_staged8.json 2>> $OUT/r2f_bench_c3.err
python - <<'PY'
import json
for t in ("bps6", "bps4", "bps3", "bps2", "staged8"):
    try:
        d = json.load(open(f"gpurun_out/r2f_bench_c3_{t}.json"))
        print(t, round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items()}, round(d["roofline"]["frac"], 4))
    except Exception as e:
        print(t, "failed", e)
PY
