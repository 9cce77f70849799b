#!/bin/bash
# decode / scan overlap: parity tests, then the headline bench with and without it
TAG=${1:-ov1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
( timeout 600 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
grep -a "passed\|failed\|rror" $OUT/pytest_gpu.log | tail -n 5
for V in overlap plain; do
  NO=0; [ "$V" = plain ] && NO=1
  DFDB_NO_OVERLAP=$NO timeout 300 python bench.py --steps 8 --warmup 3 --no-e2e > $OUT/bench_$V.json 2> $OUT/bench_$V.err
  python - <<PY
import json
try:
    d = json.load(open("$OUT/bench_$V.json"))
    print("$V", "ms/step", round(d["ms_per_step"], 3), "Grows/s", round(d["value"] / 1e9, 2), d["phases_ms_per_step"], "frac", round(d["roofline"]["frac"], 4), d["verified"]["ok"] if d["verified"] else None, "launches", d["gpu_launches"])
except Exception as e:
    print("$V", "failed", e)
PY
done
