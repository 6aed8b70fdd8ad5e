#!/usr/bin/env python
"""Per-CUDA-source-line executed warp instructions from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
fname, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":   # a CUDA line (aggregated over its SASS)
        try:
            n = int(r[hdr.index("Instructions Executed")].replace(",", ""))
        except ValueError:
            continue
        out.append((n, fname, r[0], r[1][:100]))
tot = sum(o[0] for o in out)
print("total warp instructions:", tot)
for n, f, l, sx in sorted(out, reverse=True)[:top]:
    print(f"{100 * n / tot:5.1f}%  {f}:{l}  {sx}")
