#!/bin/bash
# Whole GPU parity suite (+ optional short bench).
TAG=${1:-tests}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 1500 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -15 $OUT/pytest_gpu.log
if [ "$2" = "bench" ]; then
( timeout 900 python bench.py --steps 5 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
tail -3 $OUT/bench.err; cat $OUT/bench.json
fi
