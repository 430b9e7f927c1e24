#!/bin/bash
# End-of-round evidence: launch list of the bench command, full captures of the two hot kernels, DRAM traffic
# of the blind rotation at the bench batch size.   gpurun --timeout 1500 -- bash tools/gpu_final.sh r01
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
SMALL="python bench.py --steps 2 --warmup 3 --batch 592 --no-cpu-baseline --no-match"
timeout 300 $SMALL > $OUT/plain_$TAG.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_$TAG.csv $SMALL > $OUT/ncu_list_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:blind_rotate -s 3 -c 1 -o $OUT/prof_br_$TAG -f $SMALL > $OUT/ncu_br_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ks_gemm -s 3 -c 1 -o $OUT/prof_ks_$TAG -f $SMALL > $OUT/ncu_ks_$TAG.log 2>&1
BIG="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-match"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"blind_rotate|ks_gemm|ks_decompose" -s 9 -c 3 --csv --log-file $OUT/traffic_$TAG.csv $BIG > $OUT/ncu_traffic_$TAG.log 2>&1
tail -4 $OUT/traffic_$TAG.csv | cut -c1-400
ls -la $OUT | tail -12
