#!/bin/bash
# usage: gpu_prof.sh <tag> <kernel-regex> [rows]   -- lz4 tests, short bench, ncu --set full of one kernel
TAG=${1:-p}; KRE=${2:-lz4_decode}; ROWS=${3:-200000000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lz4" ) > $OUT/pytest_lz4.log 2>&1
tail -4 $OUT/pytest_lz4.log
if ! grep -q " passed" $OUT/pytest_lz4.log || grep -q "failed\|error" $OUT/pytest_lz4.log; then echo "LZ4 TESTS FAILED"; exit 1; fi
( timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e ) > $OUT/bench.json 2> $OUT/bench.err
tail -3 $OUT/bench.err; python -c "
import json,sys
d=json.load(open('$OUT/bench.json'))
print('ms/step',d['ms_per_step'],'decode',d['phases_ms_per_step'],'roofline',d['roofline']['frac'],'ok',d['verified'])
"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -o $OUT/prof -f \
    python bench.py --rows $ROWS --steps 2 --warmup 1 --no-e2e --no-verify > $OUT/prof.log 2>&1
ls -la $OUT
