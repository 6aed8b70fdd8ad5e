#!/bin/bash
# round 2, GPU call 24: sub-warp scatter of the tiny class with register-cached weights — A/B on the tiny workload
# (committed kernel / new at 2 CTAs per SM / new at 3 CTAs per SM with spills), 2D parity tests on the new default,
# and the HEALPix c4s sample with the asin-class pass A at 3 CTAs per SM
mkdir -p gpurun_out
B="python bench.py --extra none --no-parity --no-cpu-baseline --no-e2e"
cp sphtogrid.jl_b200/libsphtogrid_cuda.so /tmp/main.so
timeout 600 $B --workload tiny --steps 3 --warmup 2 > gpurun_out/r2w_tiny_new.json 2> gpurun_out/r2w_tiny_new.err
timeout 600 $B --workload c4s --steps 3 --warmup 1 > gpurun_out/r2w_c4s_new.json 2> gpurun_out/r2w_c4s_new.err
timeout 900 python -m pytest tests -q -m gpu -x -k "2d or 2D or golden or tiny or baseline or fp32 or healpix" > gpurun_out/r2w_tests_new.log 2>&1; tail -n 2 gpurun_out/r2w_tests_new.log
for v in old minb3; do
  cp sphtogrid.jl_b200/libs2g_alt_$v.so sphtogrid.jl_b200/libsphtogrid_cuda.so
  timeout 600 $B --workload tiny --steps 3 --warmup 2 > gpurun_out/r2w_tiny_$v.json 2> gpurun_out/r2w_tiny_$v.err
done
cp /tmp/main.so sphtogrid.jl_b200/libsphtogrid_cuda.so
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2w_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()})
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
