#!/usr/bin/env python
"""Per-CUDA-line table from an `ncu --page source --csv --print-source cuda,sass` export: warp instructions per batch, stall samples
and the top two stall reasons.  usage: ncu_by_line.py <source.csv> <file.cu> <batches> [min_instr] > profiles/<name>.txt"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
srcpath = sys.argv[2]
batches = float(sys.argv[3])
min_ins = float(sys.argv[4]) if len(sys.argv) > 4 else 10e6
src = open(srcpath).read().split('\n')
stalls = ['stall_long_sb', 'stall_short_sb', 'stall_wait', 'stall_math', 'stall_branch_resolving', 'stall_not_selected', 'stall_no_inst',
          'stall_lg', 'stall_mio', 'stall_dispatch']
cur = None; hdr = None; tot = 0; tots = 0; ci = {}
print("line  instr/batch  samples  top stalls                            source")
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; ci = {h: i for i, h in enumerate(hdr)}; continue
    if cur and r[0].isdigit() and hdr and len(r) == len(hdr):
        ln = int(r[0]); ins = int(r[ci['Instructions Executed']]); smp = int(r[ci['# Samples']])
        tot += ins; tots += smp
        if cur.endswith(srcpath.split('/')[-1]) and (ins > min_ins or smp > 2500):
            st = sorted(((int(r[ci[s]]), s.replace('stall_', '')) for s in stalls if s in ci and r[ci[s]].isdigit()), reverse=True)[:2]
            print(f"{ln:4d} {ins / batches:10.2f} {smp:9d}  {' '.join(f'{n}:{v}' for v, n in st):36s} {src[ln - 1].strip()[:100]}")
print(f"(source-level instructions in all files {tot / 1e9:.3f} G = {tot / batches:.1f} per batch, samples {tots})")
