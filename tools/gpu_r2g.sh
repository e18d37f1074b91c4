#!/bin/bash
# round 2, GPU call G: the whole -m gpu suite (no -x: every failure at once)
OUT=gpurun_out; mkdir -p $OUT
timeout 3000 python -m pytest tests -q -m gpu --durations=12 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]" | tail -60 > $OUT/r2g_pytest.log; cat $OUT/r2g_pytest.log
