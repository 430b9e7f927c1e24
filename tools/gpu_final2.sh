#!/bin/bash
# End-of-round evidence for the match path: launch list of one 64-char and one 256-char match, full capture of the
# latency blind rotation, final bench line.   gpurun --timeout 1500 -- bash tools/gpu_final2.sh r01
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
M="python tools/match_once.py 64 /a+b?c/ 256 /a+b?c/"
timeout 300 $M > $OUT/match_plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/match_launches_$TAG.csv $M > $OUT/match_ncu_list_$TAG.log 2>&1
cat $OUT/match_plain_$TAG.log
SMALL="python bench.py --steps 2 --warmup 3 --batch 148 --no-cpu-baseline --no-match"
timeout 300 $SMALL > $OUT/plain_wide_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_wide -s 3 -c 1 -o $OUT/prof_wide_$TAG -f $SMALL > $OUT/ncu_wide_$TAG.log 2>&1
tail -2 $OUT/ncu_wide_$TAG.log | cut -c1-200
timeout 900 python bench.py > $OUT/bench_final_$TAG.json 2> $OUT/bench_final_$TAG.err; echo "bench exit $?"
cut -c1-300 $OUT/bench_final_$TAG.json
