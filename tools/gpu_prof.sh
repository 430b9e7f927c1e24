#!/bin/bash
# ncu full capture of one kernel on a one-wave batch.  gpurun --timeout 900 -- bash tools/gpu_prof.sh tag [kernel-regex] [batch]
TAG=${1:-prof}
KRE=${2:-blind_rotate}
B=${3:-444}
EXTRA=${4:-}   # e.g. "--option br_samples=6"
OUT=gpurun_out
mkdir -p $OUT
SMALL="python bench.py --steps 2 --warmup 3 --batch $B --no-cpu-baseline --no-match $EXTRA"
timeout 300 $SMALL > $OUT/plain_$TAG.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -o $OUT/prof_$TAG -f $SMALL > $OUT/ncu_full_$TAG.log 2>&1
tail -3 $OUT/plain_$TAG.log | cut -c1-600; tail -5 $OUT/ncu_full_$TAG.log | cut -c1-300
