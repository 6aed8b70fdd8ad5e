#!/bin/bash
# round 2, GPU call 15: 3D scatter through shared-memory tiles (8^3 blocks, 24^3 tile), slice-size sensitivity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_3d_healpix.py tests/test_golden_vectors.py tests/test_gpu_baseline_streams.py -q -m gpu -x -k "3d or golden or 3D" > gpurun_out/r2o_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2o_tests.log; tail -n 3 gpurun_out/r2o_tests.log
B="python bench.py --extra none --no-parity --no-cpu-baseline --no-e2e"
timeout 600 $B --workload c3s --steps 3 --warmup 2 > gpurun_out/r2o_c3s.json 2> gpurun_out/r2o_c3s.err
timeout 600 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2o_c3.json 2> gpurun_out/r2o_c3.err
S2G_BATCH_PARTICLES=67108864 timeout 600 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2o_c3_b64m.json 2> gpurun_out/r2o_c3_b64m.err
S2G_BATCH_PARTICLES=16777216 timeout 600 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2o_c3_b16m.json 2> gpurun_out/r2o_c3_b16m.err
S2G_3D_TILE=0 timeout 600 $B --workload c3 --steps 3 --warmup 2 > gpurun_out/r2o_c3_notile.json 2> gpurun_out/r2o_c3_notile.err
timeout 600 $B --workload c3big --steps 3 --warmup 2 > gpurun_out/r2o_c3big.json 2> gpurun_out/r2o_c3big.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2o_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()})
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-600:])
PY
