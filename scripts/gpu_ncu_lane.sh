#!/bin/bash
# ncu --set full capture of the lane-per-block decoder on column ia (rand 1..100 Int64).
TAG=${1:-ncu_lane}
ROWS=${2:-1000000000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
DFDB_LZ4_FLAVOUR=3 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:lz4_decode_lane -s 1 -c 1 -o $OUT/lane \
   python scripts/decode_kinds.py --rows $ROWS --reps 1 --cols ia --check-rows 65536 > $OUT/ncu.log 2>&1
tail -5 $OUT/ncu.log
ls -la $OUT
