#!/bin/bash
# round 2, GPU call 3: sub-warp scatter tests + A/B, ncu captures of the HEALPix gather and its pass A, e2e phases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_2d.py tests/test_golden_vectors.py tests/test_fp32_accumulate.py -q -m gpu -x > gpurun_out/r2c_2d_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2c_2d_tests.log; tail -n 3 gpurun_out/r2c_2d_tests.log
B="python bench.py --extra none --no-parity --no-cpu-baseline"
for v in 64 0; do
  S2G_TINY_MAX_PIXELS=$v timeout 600 $B --workload tiny --steps 3 --warmup 2 --no-e2e > gpurun_out/r2c_tiny_max$v.json 2> gpurun_out/r2c_tiny_max$v.err
done
timeout 900 $B --workload c2 --steps 2 --warmup 2 > gpurun_out/r2c_c2.json 2> gpurun_out/r2c_c2.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hp_gather -c 1 -o gpurun_out/r2_prof_hpgather -f $B --workload c4t --steps 1 --warmup 0 --no-e2e > gpurun_out/r2c_ncu_hpgather.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_healpix -c 1 -o gpurun_out/r2_prof_hprec -f $B --workload c4t --steps 1 --warmup 0 --no-e2e > gpurun_out/r2c_ncu_hprec.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2c_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()}, "e2e", e.get("ms_per_step"), e.get("pinned_ms_per_step"), e.get("phases_last_call"))
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
ls -la gpurun_out/*.ncu-rep
