#!/bin/bash
# Round-2 profiles: `ncu --set full` of every kernel family (text pages only travel back), racecheck, SASS opcode histogram.
OUT=gpurun_out/${1:-prof_r2}
mkdir -p $OUT
cap() {  # cap <name> <kernel regex> <skip> <workload> [rows]
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o $OUT/$1 python scripts/prof_workloads.py $4 $5 > $OUT/$1.log 2>&1
  tail -1 $OUT/$1.log
  # gpurun brings back at most 64 MiB: keep the text pages, drop the report
  ncu -i $OUT/$1.ncu-rep --page details > $OUT/$1.details.txt 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1.raw.csv 2>/dev/null
  rm -f $OUT/$1.ncu-rep
  grep -E "^  [a-z_:A-Z<>, 0-9()*]+\(|Duration|Executed Ipc Active|DRAM Throughput|Achieved Occupancy|Registers Per" $OUT/$1.details.txt | head -8
}
cap lz4_decode_v3_strings_200M lz4_decode_v3 1 strings 200000000
cap vm_mask_strings_200M vm_mask 1 strings 200000000
cap gather_strings_200M gather_strings 1 strings 200000000
cap gather_fixed_200M gather_fixed 3 missings 200000000
cap fused_scan_mask_200M fused_scan_kernel 1 missings 200000000
cap range_stage_200M range_stage 1 range 200000000
cap agg_vm_200M agg_vm 1 arith 200000000
cap group_reduce_200M group_reduce 1 group 200000000
cap zone_map_200M zone_map 0 zone 200000000
cap lz4_compress_50M lz4_compress 1 write 50000000
# racecheck over the GPU parity tests of the scan / gather / aggregate / write kernels; the walker / consumer decoder hands data between
# warps through shared-memory rings with release / acquire flags, which racecheck reports by design: it is excluded by name
( timeout 900 compute-sanitizer --tool racecheck --kernel-name-exclude kns=lz4_decode_v3 \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "plans_match_oracle or aggregates_match or groupreduce or zone_maps_prune or flat_strings or missings or write_path or selection_stages" ) > $OUT/racecheck.txt 2>&1
tail -5 $OUT/racecheck.txt
ls -la $OUT | head -40
