#!/bin/bash
# decode + predicate + aggregate in one kernel: parity tests, then the bench with and without it.
TAG=${1:-fused}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or overlap or residency or synthetic_agg or zone" ) > $OUT/pytest.log 2>&1
tail -5 $OUT/pytest.log
if grep -q "failed\|error\|Timeout" $OUT/pytest.log; then echo "FUSED TESTS FAILED"; grep -E "Error|assert" $OUT/pytest.log | head -20; exit 1; fi
for V in "DFDB_NO_DECODE_FUSED=0" "DFDB_NO_DECODE_FUSED=1"; do
( env $V timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-variants ) > $OUT/bench_$V.json 2> $OUT/bench_$V.err
python - <<PY
import json
b=json.loads(open("$OUT/bench_$V.json").read().strip().splitlines()[-1])
print("$V", "value", round(b["value"]/1e9,2), "G rows/s  ms", round(b["ms_per_step"],3), "phases", b["phases_ms_per_step"], "verified", b["verified"]["ok"])
PY
done
