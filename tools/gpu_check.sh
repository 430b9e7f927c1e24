#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list, ncu full capture of the blind-rotate kernel.
#   gpurun --timeout 1500 -- bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
set -x
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
SMALL="python bench.py --steps 2 --warmup 3 --batch 444 --no-cpu-baseline --no-match"
timeout 300 $SMALL > $OUT/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_list_$TAG.log 2>&1
timeout 300 $SMALL > $OUT/plain2_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 3 -c 1 -o $OUT/prof_br_$TAG -f $SMALL > $OUT/ncu_full_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:keyswitch -s 3 -c 1 -o $OUT/prof_ks_$TAG -f $SMALL > $OUT/ncu_full_ks_$TAG.log 2>&1
ls -la $OUT
