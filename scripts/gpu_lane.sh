#!/bin/bash
# Lane-per-block decoder: parity tests of the codec hook, then K1 throughput per body kind against the walker / consumer flavours.
TAG=${1:-lane}
ROWS=${2:-200000000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $OUT/gpu.txt 2>&1
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lz4 and lane" ) > $OUT/pytest_lane.log 2>&1
tail -5 $OUT/pytest_lane.log
if grep -q "failed\|error\|Timeout" $OUT/pytest_lane.log; then echo "LANE TESTS FAILED"; [ "$3" = "force" ] || exit 1; fi
( DFDB_LZ4_FLAVOUR=3 timeout 900 python scripts/decode_kinds.py --rows $ROWS --reps 3 ) > $OUT/kinds_lane.txt 2> $OUT/kinds_lane.err
cut -c1-60,150-260 $OUT/kinds_lane.txt; tail -3 $OUT/kinds_lane.err
if [ "$4" != "" ]; then
( DFDB_LZ4_FLAVOUR=3 timeout 900 python scripts/decode_kinds.py --rows $4 --reps 3 --cols ia,s ) > $OUT/kinds_lane_big.txt 2> $OUT/kinds_lane_big.err
cut -c1-60,150-260 $OUT/kinds_lane_big.txt; tail -3 $OUT/kinds_lane_big.err
( timeout 900 python scripts/decode_kinds.py --rows $4 --reps 3 --cols ia,s ) > $OUT/kinds_walker_big.txt 2> $OUT/kinds_walker_big.err
cut -c1-60,150-260 $OUT/kinds_walker_big.txt
fi
