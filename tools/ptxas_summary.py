#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` logs: kernel, registers, spills, smem (skips cub kernels)."""
import glob, re, subprocess, sys
logs = sys.argv[1:] or glob.glob("sphtogrid.jl_b200/csrc/*.ptxas.log")
rows = []
for lg in logs:
    txt = open(lg).read().splitlines()
    for i, l in enumerate(txt):
        m = re.search(r"Compiling entry function '(\S+)' for", l)
        if not m:
            continue
        name = m.group(1)
        blob = " ".join(txt[i:i + 4])
        regs = re.search(r"Used (\d+) registers", blob)
        sp = re.search(r"(\d+) bytes spill stores", blob)
        sm = re.search(r"(\d+) bytes smem", blob)
        rows.append((name, int(regs.group(1)) if regs else -1, int(sp.group(1)) if sp else 0, int(sm.group(1)) if sm else 0))
names = subprocess.run(["c++filt"] + [r[0] for r in rows], capture_output=True, text=True).stdout.splitlines()
for (n, r, s, m), d in zip(rows, names):
    if "cub::" in d:
        continue
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    print(f"{d:40s} regs={r:3d} spill={s:4d} smem={m}")
