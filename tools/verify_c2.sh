#!/bin/bash
# On the GPU box: SAM byte parity at BASELINE configs[1]'s own size.  Builds the bench's 100 Mbp index and buckets
# (bench.py's prepare()), then runs the unmodified reference (`oracle/_ref/ema align -t 1`) and ema-b200 on full
# 40 000-pair buckets, plain and with -d under the pinned clock, and compares the SAM bodies.  usage: tools/verify_c2.sh <tag> [n_buckets]
TAG=${1:-r1}
NB=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
LOG=$OUT/${TAG}_verify_c2.log
: > $LOG
python - <<PY 2>&1 | tail -2 | tee -a $LOG
import sys; sys.path.insert(0, "."); sys.argv = ["bench.py"]
import bench
fa, buckets, ppb = bench.prepare("c2", "/tmp/emab_data", $NB)
print("data:", fa, len(buckets), "buckets of", ppb, "pairs")
PY
D=/tmp/emab_data/bench_c2
gcc -shared -fPIC -O2 tests/shims/faketime.c -o /tmp/faketime.so
for b in $(seq 0 $((NB-1))); do
  f=$D/ema-bin-$(printf %03d $b)
  for mode in plain d; do
    extra=""; pre=""
    if [ $mode = d ]; then extra="-d"; pre="env LD_PRELOAD=/tmp/faketime.so"; fi
    t0=$(date +%s.%N)
    $pre oracle/_ref/ema align -s $f -r $D/ref.fa -p 10x $extra -t 1 -o /tmp/ref_$b.sam 2> /tmp/ref_$b.err || echo "reference failed: $(tail -1 /tmp/ref_$b.err)" | tee -a $LOG
    t1=$(date +%s.%N)
    $pre ema_b200/ema-b200 align -s $f -r $D/ref.fa -p 10x $extra -t 16 -o /tmp/our_$b.sam 2> /tmp/our_$b.err || echo "ema-b200 failed: $(tail -1 /tmp/our_$b.err)" | tee -a $LOG
    t2=$(date +%s.%N)
    echo "bucket $b mode $mode: reference -t 1 $(python3 -c "print(round($t1 - $t0, 2))") s, ema-b200 -t 16 $(python3 -c "print(round($t2 - $t1, 2))") s (process start, index load and dense-SA build included)" | tee -a $LOG
    r=$(grep -v '^@' /tmp/ref_$b.sam | md5sum | cut -c1-32); o=$(grep -v '^@' /tmp/our_$b.sam | md5sum | cut -c1-32)
    n=$(grep -vc '^@' /tmp/ref_$b.sam)
    hr=$(grep '^@' /tmp/ref_$b.sam | grep -v '^@PG' | md5sum | cut -c1-32); ho=$(grep '^@' /tmp/our_$b.sam | grep -v '^@PG' | md5sum | cut -c1-32)
    if [ "$n" -lt 1000 ]; then v="NOT RUN (only $n records)"; elif [ "$r" = "$o" ] && [ "$hr" = "$ho" ]; then v=IDENTICAL; else v="DIFFERENT ($(diff <(grep -v '^@' /tmp/ref_$b.sam) <(grep -v '^@' /tmp/our_$b.sam) | grep -c '^<') lines)"; fi
    echo "bucket $b mode $mode: $n SAM records, body md5 ref $r ours $o, header(-@PG) $hr / $ho -> $v" | tee -a $LOG
  done
done
