#!/bin/bash
# the driver's launch line at N GPUs (one rank per GPU under torchrun); usage: tools/gpu_scale.sh <tag> <N>
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r3f}; N=${2:-4}
nvidia-smi --query-gpu=index,name --format=csv,noheader > $OUT/${TAG}_box.txt; nproc >> $OUT/${TAG}_box.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
tail -3 $OUT/${TAG}_bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
print("n", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"], 2), d["e2e"]["ms_per_step_repeats"], d["config"]["host_threads_per_rank"], d["e2e"]["stage_ms_per_step_summed_over_workers"])
PY
