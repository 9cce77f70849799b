#!/bin/bash
# One gpurun call: LZ4 + parity tests, per-kind decode throughput, a short bench (regression check of the headline).
# usage: gpurun --timeout 900 -- 'bash scripts/gpu_kinds.sh <tag> [rows]'
TAG=${1:-k1}
ROWS=${2:-200000000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( time timeout 300 python -m pytest tests -m gpu -x -q -k lz4 ) > $OUT/pytest_lz4.log 2>&1
tail -3 $OUT/pytest_lz4.log
( time timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -3 $OUT/pytest_gpu.log
( time timeout 600 python scripts/decode_kinds.py --rows $ROWS ) > $OUT/kinds.json 2> $OUT/kinds.err
tail -2 $OUT/kinds.err; cat $OUT/kinds.json
if [ "$3" != "nobench" ]; then
( time timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e ) > $OUT/bench.json 2> $OUT/bench.err
tail -2 $OUT/bench.err; python - <<PY
import json
try:
    d = json.load(open("$OUT/bench.json"))
    print("bench", d["value"], d["ms_per_step"], d["phases_ms_per_step"], d["verified"])
except Exception as e:
    print("bench parse failed", e)
PY
fi
