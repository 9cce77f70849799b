#!/bin/bash
# fused decode + scan (scan warps inside the decode kernel): parity test of the option, bench with and without
OUT=gpurun_out/${1:-fused2}
mkdir -p $OUT
( timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or residency or overlap" ) > $OUT/pytest.log 2>&1
tail -3 $OUT/pytest.log
if grep -q "failed\|error\|Timeout" $OUT/pytest.log; then grep -E "Error|assert|FAILED" $OUT/pytest.log | head; exit 1; fi
for V in "DFDB_NO_DECODE_FUSED=0" "DFDB_NO_DECODE_FUSED=1"; do
( env $V timeout 900 python bench.py --steps 10 --warmup 3 --no-e2e --no-variants ) > $OUT/bench_$V.json 2> $OUT/bench_$V.err
python - <<PY
import json
b=json.loads(open("$OUT/bench_$V.json").read().strip().splitlines()[-1])
print("$V", "value", round(b["value"]/1e9,2), "G rows/s  ms", round(b["ms_per_step"],3), "phases", b["phases_ms_per_step"], "launches", b["gpu_launches"], "verified", b["verified"]["ok"], b["verified"]["rel_err"])
PY
done
