#!/bin/bash
# one `ncu --set full` capture of a K1 kernel on one column of the decode_kinds table: gpu_ncu_kinds.sh <tag> <kernel regex> <column> [rows] [env]
OUT=gpurun_out/$1
mkdir -p $OUT
env $5 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -o $OUT/cap python scripts/decode_kinds.py --cols $3 --rows ${4:-200000000} --reps 1 > $OUT/ncu.log 2>&1
tail -2 $OUT/ncu.log
ncu -i $OUT/cap.ncu-rep --page details > $OUT/details.txt 2>/dev/null
ncu -i $OUT/cap.ncu-rep --page source --csv --print-source cuda,sass > $OUT/source.csv 2>/dev/null
rm -f $OUT/cap.ncu-rep
grep -E "Duration|Executed Ipc Active|Issue Slots Busy|Achieved Occupancy|Registers Per|Executed Instructions  |Warp Cycles Per Issued" $OUT/details.txt | head
