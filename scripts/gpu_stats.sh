#!/bin/bash
# usage: gpu_stats.sh <tag> [rows] -- bench with the decoder's diagnostics counters on (numbers are NOT bench values)
TAG=${1:-s}; ROWS=${2:-1000000000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lz4" ) > $OUT/pytest_lz4.log 2>&1
tail -3 $OUT/pytest_lz4.log
( DFDB_B200_LIB=$PWD/dataframedbs.jl_b200/lib/libdfdb_b200_stats.so DFDB_LZ4_STATS=1 timeout 150 python bench.py --rows $ROWS --steps 1 --warmup 0 --no-e2e --no-verify ) > $OUT/bench_stats.json 2> $OUT/bench_stats.err
grep "lz4 v2" $OUT/bench_stats.err $OUT/bench_stats.json | head -12
( timeout 150 python bench.py --rows $ROWS --steps 5 --warmup 3 --no-e2e ) > $OUT/bench.json 2> $OUT/bench.err
python -c "
import json
d=json.load(open('$OUT/bench.json'))
print('ms/step',d['ms_per_step'],'decode',d['phases_ms_per_step'],'roofline',d['roofline']['frac'],'ok',d['verified'])
"
