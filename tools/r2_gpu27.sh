#!/bin/bash
# round 2, GPU call 27: 2D scatter kernels with pass-A weights cached in shared memory (flattened footprints in the
# warp-per-particle kernel) — tiny workload, C2 / C5 sample unchanged?, 2D parity tests, ncu of the tiny sample
mkdir -p gpurun_out
B="python bench.py --extra none --no-parity --no-cpu-baseline --no-e2e"
timeout 600 $B --workload tiny --steps 3 --warmup 2 > gpurun_out/r2z_tiny.json 2> gpurun_out/r2z_tiny.err
timeout 600 $B --workload c2 --steps 3 --warmup 2 > gpurun_out/r2z_c2.json 2> gpurun_out/r2z_c2.err
timeout 600 $B --workload c5s --steps 3 --warmup 2 > gpurun_out/r2z_c5s.json 2> gpurun_out/r2z_c5s.err
timeout 900 python -m pytest tests -q -m gpu -x -k "2d or 2D or golden or tiny or baseline or fp32 or sedov or stokes" > gpurun_out/r2z_tests.log 2>&1; tail -n 2 gpurun_out/r2z_tests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scatter2d -s 2 -c 2 -f -o gpurun_out/r2z_k_scatter2d_tiny python bench.py --workload tinys --steps 1 --warmup 1 --extra none --no-parity --no-cpu-baseline --no-e2e > gpurun_out/r2z_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2z_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()})
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
