#!/bin/bash
# Quick GPU iteration: decode tests first (bounded), then the whole GPU suite, then a short bench.
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lz4" ) > $OUT/pytest_lz4.log 2>&1
tail -15 $OUT/pytest_lz4.log
if grep -q "failed\|error\|Timeout" $OUT/pytest_lz4.log; then echo "LZ4 TESTS FAILED"; [ "$2" = "force" ] || exit 1; fi
( timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -5 $OUT/pytest_gpu.log
( timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e ) > $OUT/bench.json 2> $OUT/bench.err
tail -3 $OUT/bench.err; cat $OUT/bench.json
