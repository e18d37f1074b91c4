#!/bin/bash
# round 2, GPU call N (2 GPUs): the driver's launch lines: reference arm, N=1, then N=2 under torchrun; multi-device tests
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/r2n_box.txt; nproc >> $OUT/r2n_box.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 8 --warmup 2 > $OUT/r2n_bench_reference.json 2> $OUT/r2n_bench_reference.err; tail -2 $OUT/r2n_bench_reference.err; cut -c1-600 $OUT/r2n_bench_reference.json
timeout 900 python bench.py --gpus 1 --steps 16 --warmup 3 > $OUT/r2n_bench_n1.json 2> $OUT/r2n_bench_n1.err; tail -2 $OUT/r2n_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 16 --warmup 3 > $OUT/r2n_bench_n2.json 2> $OUT/r2n_bench_n2.err; tail -3 $OUT/r2n_bench_n2.err
python - <<'PY'
import json
for t in ("n1", "n2"):
    try:
        d = json.loads(open(f"gpurun_out/r2n_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(t, "failed", e)
PY
timeout 900 python -m pytest tests/test_gpu_sam.py -x -q -m gpu -k "multi_device or single_bucket" 2>&1 | tail -3
