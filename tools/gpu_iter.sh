#!/bin/bash
# quick iteration: smoke, GPU parity tests, short bench line.   gpurun --timeout 900 -- bash tools/gpu_iter.sh tag
TAG=${1:-iter}
OUT=gpurun_out
mkdir -p $OUT
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -20 $OUT/smoke_$TAG.log; exit 1; }
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
tail -15 $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
