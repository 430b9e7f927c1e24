#!/bin/bash
# Last evidence run of round 2 (new default throughput kernel: I2F digits, a plane per component):
# bench line, launch list, full capture of the blind rotation, DRAM traffic at the bench batch.
#   gpurun --timeout 1500 -- bash tools/gpu_r02_last.sh r02c
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cut -c1-700 $OUT/bench_$TAG.json
SMALL="python bench.py --steps 2 --warmup 3 --batch 592 --no-cpu-baseline --no-match"
timeout 300 $SMALL > $OUT/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_list_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate_fused -s 3 -c 1 -o $OUT/prof_br_$TAG -f $SMALL > $OUT/ncu_br_$TAG.log 2>&1
BIG="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-match"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"blind_rotate|ks_umma|ks_decompose" -s 9 -c 3 --csv --log-file $OUT/traffic_$TAG.csv $BIG > $OUT/ncu_traffic_$TAG.log 2>&1
tail -4 $OUT/traffic_$TAG.csv | cut -c1-400
