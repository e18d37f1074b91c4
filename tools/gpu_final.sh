#!/bin/bash
# round-end evidence on one B200: full gpu suite, smoke, the driver's bench lines (ours + reference arm), launch list, ncu --set full of one bucket
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r4a}
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader > $OUT/${TAG}_box.txt; nproc >> $OUT/${TAG}_box.txt; lscpu | grep "Model name" >> $OUT/${TAG}_box.txt
timeout 2400 python -m pytest tests -x -q -m gpu 2>&1 | grep -v "^BWA\|^Processing\|M::bwa\|^\[index\]\|^\[bwa\|^\[bwt\|^\[main\]" | tail -30 > $OUT/${TAG}_pytest_gpu.log; tail -3 $OUT/${TAG}_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; tail -2 $OUT/${TAG}_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; cut -c1-300 $OUT/${TAG}_bench_reference.json
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $OUT/${TAG}_clocks.csv &
SMI=$!
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; tail -2 $OUT/${TAG}_bench_n1.err
kill $SMI
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step_repeats"], {k: round(v, 3) for k, v in d["device_ms_per_step"].items()}, "roof", round(d["roofline"]["frac"], 3), d["cpu_baseline"]["value"], d["sw_microbench"]["roofline"]["frac"])
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches_c3.csv \
    python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --single-only > $OUT/${TAG}_ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_seed_rq|k_chain|k_ext_wave|k_align1_replay|k_rescue_sw|k_rescue$|k_glob_wave|k_glob_wide|k_finalize|k_em' -s 22 -c 11 \
    -f -o $OUT/${TAG}_prof_c3 python bench.py --workload c3 --steps 1 --warmup 2 --single-only --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
tail -1 $OUT/${TAG}_ncu_full.log | cut -c1-200
ls -la $OUT/${TAG}_prof_c3.ncu-rep
