#!/usr/bin/env python
"""sass_summary.py -- opcode histogram per kernel of the built library (cuobjdump -sass), written to profiles/.

What to look for (B200_PROFILING.md): UBLKCP = bulk copy / TMA engine (cp.async.bulk), LDGSTS = cp.async, SYNCS.* = mbarrier,
no HMMA / UTC*MMA anywhere (nothing on this path is a contraction), and the size of each kernel's code against the 32 KB
instruction cache that matters for the kernels that run one warp per scheduler."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dataframedbs.jl_b200", "lib", "libdfdb_b200.so")
KEY = ("UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "LDGDEPBAR", "LDS", "STS", "LDG", "STG", "ATOMG", "ATOMS", "RED", "SHFL", "VOTE", "MATCH", "BAR", "HMMA", "UTC")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    lines = [f"SASS summary of {os.path.relpath(LIB, ROOT)} (sm_100a), one line per kernel: instructions, code bytes, then selected opcodes", ""]
    total = collections.Counter()
    for name, c in kernels.items():
        n = sum(c.values())
        sel = " ".join(f"{k}={sum(v for op, v in c.items() if op.startswith(k))}" for k in KEY if any(op.startswith(k) for op in c))
        top = " ".join(f"{op}:{v}" for op, v in c.most_common(6))
        lines.append(f"{name}\n    {n} instr ({n * 16} B)  {sel}\n    top: {top}")
        total.update(c)
    lines += ["", "whole library: " + " ".join(f"{k}={sum(v for op, v in total.items() if op.startswith(k))}" for k in KEY)]
    text = "\n".join(lines) + "\n"
    dst = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_summary.txt")
    open(dst, "w").write(text)
    print(text[-600:])


if __name__ == "__main__":
    main()
