#!/bin/bash
# bench.py as the driver runs it (both arms), N = 1.
TAG=${1:-bench}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python bench.py --impl reference --gpus 1 --steps ${2:-5} --warmup ${3:-2} ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
tail -2 $OUT/bench_ref.err; cut -c1-400 $OUT/bench_ref.json
( timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $OUT/bench.json 2> $OUT/bench.err
tail -5 $OUT/bench.err; cat $OUT/bench.json
