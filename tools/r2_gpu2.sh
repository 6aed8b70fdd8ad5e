#!/bin/bash
# round 2, GPU call 2: tile-gather tests, full suite, A/B benches (HEALPix gather, 3D cell list, overlapped staging)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_3d_healpix.py -q -m gpu -x -k "tile_gather or deposit_3d or sphmapping_3d" --durations=5 > gpurun_out/r2b_new_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_new_tests.log; tail -n 4 gpurun_out/r2b_new_tests.log
timeout 1200 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2b_all_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2b_all_tests.log; tail -n 4 gpurun_out/r2b_all_tests.log
B="python bench.py --extra none --no-parity --no-cpu-baseline"
for v in 1 0; do
  S2G_HP_GATHER=$v timeout 600 $B --workload c4s --steps 2 --warmup 1 --no-e2e > gpurun_out/r2b_c4s_gather$v.json 2> gpurun_out/r2b_c4s_gather$v.err
  S2G_3D_CACHE=$v timeout 600 $B --workload c3s --steps 3 --warmup 2 --no-e2e > gpurun_out/r2b_c3s_cache$v.json 2> gpurun_out/r2b_c3s_cache$v.err
done
for g in 6 12 24; do
  S2G_HP_GATHER_MIN_PIXELS=$g timeout 600 $B --workload c4s --steps 2 --warmup 1 --no-e2e > gpurun_out/r2b_c4s_min$g.json 2> gpurun_out/r2b_c4s_min$g.err
done
timeout 900 $B --workload c2 --steps 3 --warmup 2 > gpurun_out/r2b_c2.json 2> gpurun_out/r2b_c2.err
S2G_STAGE_OVERLAP=0 timeout 900 $B --workload c2 --steps 3 --warmup 2 > gpurun_out/r2b_c2_nooverlap.json 2> gpurun_out/r2b_c2_nooverlap.err
timeout 900 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2b_c3.json 2> gpurun_out/r2b_c3.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2b_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()}, "pairs",d["config"]["pairs"], "e2e", e.get("ms_per_step"), e.get("pinned_ms_per_step"))
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
