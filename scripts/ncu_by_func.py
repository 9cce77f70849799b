#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` export per source function: samples, instructions,
stall reasons.  usage: ncu_by_func.py <export.csv> <source.cu> [batches]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
srcfile = sys.argv[2]
nb = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
cur = None; hdr = None; data = collections.defaultdict(list)
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr and cur and r[0] != "" and len(r) == len(hdr): data[cur].append(r)
ci = {h: i for i, h in enumerate(hdr)}
stalls = ['stall_branch_resolving', 'stall_lg', 'stall_long_sb', 'stall_math', 'stall_mio', 'stall_no_inst', 'stall_not_selected',
          'stall_selected', 'stall_short_sb', 'stall_sleep', 'stall_wait', 'stall_barrier', 'stall_dispatch', 'stall_membar']
src = open(srcfile).read().split('\n')
funcs = [(i + 1, l) for i, l in enumerate(src) if re.match(r'^(static )?__device__|^__global__', l)]
funcs.append((len(src) + 1, 'END'))
ts = sum(int(r[ci['# Samples']]) for v in data.values() for r in v)
ti = sum(int(r[ci['Instructions Executed']]) for v in data.values() for r in v)
print(f"samples {ts} instructions {ti} ({ti / nb:.1f} per batch)")
print("%-24s %6s %6s %8s " % ("func", "smp%", "ins%", "ins/b") + " ".join("%7s" % s.replace('stall_', '')[:7] for s in stalls))
def row(name, rs):
    c = collections.Counter(); s_ = i_ = 0
    for r in rs:
        s_ += int(r[ci['# Samples']]); i_ += int(r[ci['Instructions Executed']])
        for s in stalls:
            try: c[s] += int(r[ci[s]])
            except Exception: pass
    if s_ > ts * 0.002 or i_ > ti * 0.002:
        print("%-24s %6.1f %6.1f %8.1f " % (name[:24], 100 * s_ / ts, 100 * i_ / ti, i_ / nb) + " ".join("%7d" % c[s] for s in stalls))
f = [f for f in data if f.endswith(srcfile.split('/')[-1])]
if f:
    for (a, name), (b, _) in zip(funcs, funcs[1:]):
        m = re.search(r'(\w+)\(', name)
        row(m.group(1) if m else name, [r for r in data[f[0]] if a <= int(r[0]) < b])
for f2 in data:
    if not f or f2 != f[0]: row(f2.split('/')[-1], data[f2])
