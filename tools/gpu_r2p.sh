#!/bin/bash
# round 2, GPU call P: the one-hot / k-mer-table / unique-locus seeding form (seed_hot.cuh): parity, then c3 bench per variant
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]" | tail -15 > $OUT/r2p_pytest.log; cat $OUT/r2p_pytest.log
B="python bench.py --workload c3 --steps 16 --warmup 3 --no-cpu-baseline"
run() { tag=$1; shift; env "$@" timeout 600 $B > $OUT/r2p_bench_c3_$tag.json 2>> $OUT/r2p_bench_c3.err; }
run hot6 EMAB_SEED_MODE=5
run hot8 EMAB_SEED_MODE=5 EMAB_SEED_BPS=8
run hot4 EMAB_SEED_MODE=5 EMAB_SEED_BPS=4
run hot6k12 EMAB_SEED_MODE=5 EMAB_KMER_K=12
run wide EMAB_SEED_MODE=3
grep -i "resident\|error" $OUT/r2p_bench_c3.err | tail -8
python - <<'PY'
import json
for t in ("hot6", "hot8", "hot4", "hot6k12", "wide"):
    try:
        d = json.load(open(f"gpurun_out/r2p_bench_c3_{t}.json"))
        print(t, round(d["value"]), round(d["e2e"]["value"]), {k: round(v, 3) for k, v in d["device_ms_per_step"].items()}, d["roofline"]["algorithmic_bytes_per_launch"])
    except Exception as e:
        print(t, "failed", e)
PY
