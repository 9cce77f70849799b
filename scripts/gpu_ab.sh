#!/bin/bash
# A/B of decoder variants built as lib/libdfdb_b200_<V>.so: the headline bench's decode phase for each.
# usage: gpurun --timeout 900 -- 'bash scripts/gpu_ab.sh <tag> V1 V2 ...'   ("main" = the product library)
TAG=$1; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for V in "$@"; do
  LIB=$PWD/dataframedbs.jl_b200/lib/libdfdb_b200_$V.so
  [ "$V" = main ] && LIB=$PWD/dataframedbs.jl_b200/lib/libdfdb_b200.so
  DFDB_B200_LIB=$LIB timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e --no-verify > $OUT/bench_$V.json 2> $OUT/bench_$V.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$V.json"))
    print("$V", "ms/step", round(d["ms_per_step"], 3), d["phases_ms_per_step"])
except Exception as e:
    print("$V", "failed", e)
PY
done
