#!/bin/bash
# full GPU suite + smoke; usage: tools/gpu_full.sh <tag>
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-full}
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]\|^\[bwa\|^\[bwt\|^\[main\]" | tail -30 > $OUT/${TAG}_pytest_gpu.log; tail -5 $OUT/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -3 $OUT/${TAG}_smoke.log
