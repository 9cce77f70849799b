#!/bin/bash
# closing captures of round 2: the kernels that changed after the profile sweep (group_reduce, byte-stream and long-sequence decoders)
OUT=gpurun_out/${1:-caps_final}
mkdir -p $OUT
cap() {  # cap <name> <kernel regex> <skip> <command...>
  local name=$1 re=$2 skip=$3; shift 3
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -o $OUT/$name "$@" > $OUT/$name.log 2>&1
  ncu -i $OUT/$name.ncu-rep --page details > $OUT/$name.details.txt 2>/dev/null
  ncu -i $OUT/$name.ncu-rep --page raw --csv > $OUT/$name.raw.csv 2>/dev/null
  rm -f $OUT/$name.ncu-rep
  grep -E "Duration|Executed Ipc Active|DRAM Throughput|Registers Per" $OUT/$name.details.txt | head -4
}
cap group_reduce_200M group_reduce 1 python scripts/prof_workloads.py group 200000000
cap lz4_decode_bytes_strings_200M lz4_decode_bytes 1 python scripts/decode_kinds.py --cols s --rows 200000000 --reps 1
cap lz4_decode_long_missingf64_200M lz4_decode_long 1 python scripts/decode_kinds.py --cols mb --rows 200000000 --reps 1
python scripts/group_time.py 200000000 2>&1 | tail -6
