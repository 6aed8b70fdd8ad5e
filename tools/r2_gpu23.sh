#!/bin/bash
# round 2, GPU call 23: 3D scatter A/B on C3 — committed kernel vs (key-ordered class lists + 32-bit linear cell offsets
# + 2 CTAs/SM pinned) vs the same with the grouped conversion / edge selects; 3D parity tests on the new default
mkdir -p gpurun_out
B="python bench.py --extra none --no-parity --no-cpu-baseline --no-e2e --workload c3 --steps 3 --warmup 1"
cp sphtogrid.jl_b200/libsphtogrid_cuda.so /tmp/main.so
timeout 600 $B > gpurun_out/r2v_c3_new.json 2> gpurun_out/r2v_c3_new.err
S2G_3D_ORDER=0 timeout 600 $B > gpurun_out/r2v_c3_new_noorder.json 2> gpurun_out/r2v_c3_new_noorder.err
timeout 900 python -m pytest tests -q -m gpu -x -k "3d or 3D or c3 or C3 or golden or staging or sedov" > gpurun_out/r2v_tests_new.log 2>&1; tail -n 2 gpurun_out/r2v_tests_new.log
for v in old gc; do
  cp sphtogrid.jl_b200/libs2g_alt_$v.so sphtogrid.jl_b200/libsphtogrid_cuda.so
  timeout 600 $B > gpurun_out/r2v_c3_$v.json 2> gpurun_out/r2v_c3_$v.err
done
timeout 900 python -m pytest tests -q -m gpu -x -k "3d or 3D or c3 or C3 or golden" > gpurun_out/r2v_tests_gc.log 2>&1; tail -n 2 gpurun_out/r2v_tests_gc.log
cp /tmp/main.so sphtogrid.jl_b200/libsphtogrid_cuda.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2v_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()})
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
