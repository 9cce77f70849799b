#!/bin/bash
# One gpurun call: GPU parity tests, bench (both arms), ncu launch list, ncu --set full of the hot kernels.
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
# full capture of the two hot kernels of the headline bench at 1B rows
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_decode_spec -s 1 -c 1 -o $OUT/prof_lz4 -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-verify > $OUT/prof_lz4.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_scan -s 1 -c 1 -o $OUT/prof_scan -f \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-verify > $OUT/prof_scan.log 2>&1
# the general decoder and the string gather on configs[2] (200M-row string table)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:lz4_decode_v3 -s 1 -c 1 -o $OUT/prof_lz4_v3 -f \
    python bench_configs.py --config 3 --reps 1 > $OUT/prof_lz4_v3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gather_strings -s 1 -c 1 -o $OUT/prof_gather_str -f \
    python bench_configs.py --config 3 --reps 1 > $OUT/prof_gather_str.log 2>&1
timeout 600 python scripts/decode_kinds.py --rows 200000000 > $OUT/kinds.json 2> $OUT/kinds.err
# configs[0] (the reference's own CPU-runnable case): 10M rows, same query, for the record
timeout 300 python bench.py --rows 10000000 --steps 20 --warmup 3 --no-e2e > $OUT/bench_10M.json 2> $OUT/bench_10M.err
timeout 400 python bench_configs.py --config 3 > $OUT/c3.json 2> $OUT/c3.err
timeout 400 python bench_configs.py --config 4 > $OUT/c4.json 2> $OUT/c4.err
ls -la $OUT
