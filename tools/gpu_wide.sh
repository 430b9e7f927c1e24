#!/bin/bash
# latency-variant iteration: parity tests of both blind rotations, launch time versus batch size, short bench.
#   gpurun --timeout 1200 -- bash tools/gpu_wide.sh tag
TAG=${1:-wide}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants_agree or trivial_inputs or stagewise or boundaries" > $OUT/pytest_wide_$TAG.log 2>&1
echo "pytest(wide) exit $?" | tee -a $OUT/pytest_wide_$TAG.log
tail -8 $OUT/pytest_wide_$TAG.log
PROBE_VARIANT=cluster timeout 120 python tools/wide_probe.py 1 37 72 74 75 148 2>&1 | grep "^cluster"
timeout 300 python tools/wave_times.py $OUT/wave_times_$TAG.json > $OUT/wave_times_$TAG.log 2>&1; echo "wave exit $?"
cat $OUT/wave_times_$TAG.log
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_$TAG.log
tail -6 $OUT/pytest_$TAG.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench exit $?"
cat $OUT/bench_$TAG.json; tail -5 $OUT/bench_$TAG.err
