#!/bin/bash
# round 2, GPU call 14: A/B of k_scatter3d at 3 CTAs/SM (80 registers, 640-entry cell lists) against the default (2 CTAs/SM)
mkdir -p gpurun_out
B="python bench.py --extra none --no-parity --no-cpu-baseline --no-e2e"
cp sphtogrid.jl_b200/libsphtogrid_cuda.so /tmp/main.so
timeout 600 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2n_c3_default.json 2> gpurun_out/r2n_c3_default.err
cp sphtogrid.jl_b200/libs2g_alt_3d.so sphtogrid.jl_b200/libsphtogrid_cuda.so
timeout 600 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2n_c3_alt.json 2> gpurun_out/r2n_c3_alt.err
S2G_3D_CACHE=0 timeout 600 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2n_c3_alt_nocache.json 2> gpurun_out/r2n_c3_alt_nocache.err
timeout 300 python -m pytest tests/test_gpu_parity_3d_healpix.py -q -m gpu -x -k "3d" > gpurun_out/r2n_tests_alt.log 2>&1; tail -n 2 gpurun_out/r2n_tests_alt.log
cp /tmp/main.so sphtogrid.jl_b200/libsphtogrid_cuda.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2n_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()})
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
