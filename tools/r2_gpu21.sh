#!/bin/bash
# round 2, GPU call 21: A/B of the rim-factor skip in the HEALPix gather kernels (alt library) on c4s and the full C4
mkdir -p gpurun_out
B="python bench.py --extra none --no-parity --no-cpu-baseline --no-e2e"
cp sphtogrid.jl_b200/libsphtogrid_cuda.so /tmp/main.so
timeout 600 $B --workload c4s --steps 3 --warmup 1 > gpurun_out/r2u_c4s_default.json 2> gpurun_out/r2u_c4s_default.err
cp sphtogrid.jl_b200/libs2g_alt_hp.so sphtogrid.jl_b200/libsphtogrid_cuda.so
timeout 600 $B --workload c4s --steps 3 --warmup 1 > gpurun_out/r2u_c4s_alt.json 2> gpurun_out/r2u_c4s_alt.err
timeout 1200 $B --workload c4 --steps 1 --warmup 1 > gpurun_out/r2u_c4_alt.json 2> gpurun_out/r2u_c4_alt.err
timeout 600 python -m pytest tests/test_gpu_parity_3d_healpix.py -q -m gpu -x -k "healpix" > gpurun_out/r2u_tests_alt.log 2>&1; tail -n 2 gpurun_out/r2u_tests_alt.log
cp /tmp/main.so sphtogrid.jl_b200/libsphtogrid_cuda.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2u_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()})
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
