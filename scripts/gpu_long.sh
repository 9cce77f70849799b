#!/bin/bash
# long-sequence decoder: codec parity tests (every flavour), K1 per kind with the flavours chosen at load and with the long decoder forced.
OUT=gpurun_out/${1:-long}
mkdir -p $OUT
( timeout 900 python -m pytest tests -m gpu -x -q -k "lz4 or codec or golden or fuzz or corrupt" ) > $OUT/pytest.log 2>&1
tail -4 $OUT/pytest.log
if grep -q "failed\|error\|Timeout" $OUT/pytest.log; then grep -E "Error|assert|FAILED" $OUT/pytest.log | head; fi
for F in ${FLAVOURS:-0 5}; do
echo "flavour $F"
( DFDB_LZ4_FLAVOUR=$F timeout 600 python scripts/decode_kinds.py --rows 200000000 --reps 3 ) > $OUT/kinds_f$F.txt 2> $OUT/kinds_f$F.err
cut -c1-60,150-260 $OUT/kinds_f$F.txt
done
