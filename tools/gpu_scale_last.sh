#!/bin/bash
# Multi-GPU evidence with the final kernels: bench line and level-sharded match records at N ranks of one box.
#   gpurun --gpus N --timeout 900 -- bash tools/gpu_scale_last.sh N tag
N=${1:-2}
TAG=${2:-r02c}
OUT=gpurun_out
mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_n${N}_$TAG.json 2> $OUT/bench_n${N}_$TAG.err; echo "bench exit $?"; cut -c1-400 $OUT/bench_n${N}_$TAG.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/dist_match_check.py > $OUT/dist_match_n${N}_$TAG.jsonl 2> $OUT/dist_match_n${N}_$TAG.err; echo "dist exit $?"; cut -c1-300 $OUT/dist_match_n${N}_$TAG.jsonl
