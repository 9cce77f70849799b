#!/bin/bash
# Round-2 closing run on one GPU: GPU parity tests, bench.py as the driver runs it (both arms), launch list of the bench step, configs[0].
TAG=${1:-r2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $OUT/pytest_gpu.log 2>&1
( timeout 300 python __graft_entry__.py smoke ) > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
tail -3 $OUT/pytest_gpu.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 ) > $OUT/bench_ref.json 2> $OUT/bench_ref.err
cut -c1-600 $OUT/bench_ref.json
( time timeout 1500 python bench.py --gpus 1 --steps 10 --warmup 3 ) > $OUT/bench.json 2> $OUT/bench.err
tail -3 $OUT/bench.err; cat $OUT/bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-verify --no-variants > $OUT/launches_bench.log 2>&1
grep -c "lz4_decode\|fused_scan\|agg_finalize" $OUT/launches.csv
timeout 300 python bench.py --rows 10000000 --steps 20 --warmup 3 --no-e2e --no-variants > $OUT/bench_10M.json 2> $OUT/bench_10M.err
cut -c1-300 $OUT/bench_10M.json
ls -la $OUT
