#!/bin/bash
# Round-2 profiles: launch list of the bench step, `ncu --set full` of every kernel family, SASS opcode histogram.
OUT=gpurun_out/${1:-prof_r2}
mkdir -p $OUT
cap() {  # cap <name> <kernel regex> <skip> <workload> [rows]
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $OUT/$1 python scripts/prof_workloads.py $4 $5 > $OUT/$1.log 2>&1
  tail -1 $OUT/$1.log
  # gpurun brings back at most 64 MiB: keep the text pages, drop the report
  ncu -i $OUT/$1.ncu-rep --page details > $OUT/$1.details.txt 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1.raw.csv 2>/dev/null
  rm -f $OUT/$1.ncu-rep
}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-variants --no-verify > $OUT/launches_bench.log 2>&1
cap lz4_decode_v2_1B lz4_decode_v2 2 walker 1000000000
cap lz4_decode_lane_1B lz4_decode_lane 2 lane 1000000000
cap fused_scan_tma_1B fused_scan_tma 2 walker 1000000000
cap vm_mask_strings_200M vm_mask 1 strings 200000000
cap gather_strings_200M gather_strings 1 strings 200000000
cap gather_fixed_200M gather_fixed 3 missings 200000000
cap fused_scan_mask_200M fused_scan_kernel 1 missings 200000000
cap range_stage_200M range_stage 1 range 200000000
cap agg_vm_200M agg_vm 1 arith 200000000
cap group_reduce_200M group_reduce 1 group 200000000
cap zone_map_200M zone_map 0 zone 200000000
cap lz4_compress_50M lz4_compress 1 write 50000000
ls -la $OUT | head -40
