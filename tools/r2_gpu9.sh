#!/bin/bash
# round 2, GPU call 9: NT=5 list really used, 3D micro-optimisations, ncu --set full of the HEALPix gather passes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_3d_healpix.py tests/test_golden_vectors.py -q -m gpu -x > gpurun_out/r2i_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2i_tests.log; tail -n 3 gpurun_out/r2i_tests.log
B="python bench.py --extra none --no-parity --no-cpu-baseline"
timeout 600 $B --workload c4s --steps 2 --warmup 1 --no-e2e > gpurun_out/r2i_c4s.json 2> gpurun_out/r2i_c4s.err
timeout 1200 $B --workload c4 --steps 1 --warmup 1 --no-e2e > gpurun_out/r2i_c4.json 2> gpurun_out/r2i_c4.err
timeout 600 $B --workload c3s --steps 3 --warmup 2 --no-e2e > gpurun_out/r2i_c3s.json 2> gpurun_out/r2i_c3s.err
timeout 600 $B --workload c3 --steps 3 --warmup 2 --no-e2e > gpurun_out/r2i_c3.json 2> gpurun_out/r2i_c3.err
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_hp_gather.3, 0, 0, 5" -c 1 -o gpurun_out/r2_prof_hpgatherB -f $B --workload c4s --steps 1 --warmup 0 --no-e2e > gpurun_out/r2i_ncu_hpgB.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_hp_gather.3, 0, 1, 5" -c 1 -o gpurun_out/r2_prof_hpgatherA -f $B --workload c4s --steps 1 --warmup 0 --no-e2e > gpurun_out/r2i_ncu_hpgA.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2i_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()}, "pairs", d["config"]["pairs"])
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-400:])
PY
ls -la gpurun_out/r2_prof_hpgather[AB].ncu-rep
