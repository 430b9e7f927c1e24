#!/bin/bash
# What the driver runs at round end, in one call: smoke, GPU tests, both bench arms with default flags.
#   gpurun --timeout 1800 -- bash tools/gpu_driverlike.sh tag
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) > $OUT/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -4 $OUT/smoke_$TAG.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest_$TAG.log
( time timeout 600 python bench.py --impl reference ) > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "reference arm exit $?"; cut -c1-400 $OUT/bench_ref_$TAG.json; tail -4 $OUT/bench_ref_$TAG.err
( time timeout 900 python bench.py ) > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"; cut -c1-300 $OUT/bench_$TAG.json; tail -4 $OUT/bench_$TAG.err
