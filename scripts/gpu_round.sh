#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), ncu launch list, ncu --set full of the two hot kernels.
# usage: gpurun --timeout 1800 -- 'bash scripts/gpu_round.sh <tag>'
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
tail -5 $OUT/pytest_gpu.log
( time timeout 900 python bench.py --steps 10 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
tail -3 $OUT/bench.err; cat $OUT/bench.json
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cat $OUT/bench_ref.json
# launch list (cold-cache, serialised: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-verify > $OUT/launches_bench.log 2>&1
# full capture of the two hot kernels on a 200M-row table (>> L2), 2 launches each
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_decode_v2 -s 1 -c 1 -o $OUT/prof_lz4 -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-verify > $OUT/prof_lz4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_scan -s 1 -c 1 -o $OUT/prof_scan -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-verify > $OUT/prof_scan.log 2>&1
ls -la $OUT
