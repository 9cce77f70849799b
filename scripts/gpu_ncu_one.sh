#!/bin/bash
# one `ncu --set full` capture: gpu_ncu_one.sh <tag> <kernel regex> <skip> <workload> <rows> [env]
OUT=gpurun_out/$1
mkdir -p $OUT
env $6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $OUT/cap python scripts/prof_workloads.py $4 $5 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
ncu -i $OUT/cap.ncu-rep --page details > $OUT/details.txt 2>/dev/null
ncu -i $OUT/cap.ncu-rep --page raw --csv > $OUT/raw.csv 2>/dev/null
ncu -i $OUT/cap.ncu-rep --page source --csv > $OUT/source.csv 2>/dev/null
ls -la $OUT
