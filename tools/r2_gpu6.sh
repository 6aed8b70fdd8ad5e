#!/bin/bash
# round 2, GPU call 6: adaptive-series gather, the DEFAULT bench line (all extras), ncu --set full captures for profiles/
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_3d_healpix.py tests/test_golden_vectors.py -q -m gpu -x -k "healpix or golden" > gpurun_out/r2f_hp_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2f_hp_tests.log; tail -n 3 gpurun_out/r2f_hp_tests.log
B="python bench.py --extra none --no-parity --no-cpu-baseline"
timeout 600 $B --workload c4s --steps 2 --warmup 1 --no-e2e > gpurun_out/r2f_c4s.json 2> gpurun_out/r2f_c4s.err
timeout 1200 $B --workload c4 --steps 1 --warmup 1 --no-e2e > gpurun_out/r2f_c4.json 2> gpurun_out/r2f_c4.err
( time timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/r2f_default.json 2> gpurun_out/r2f_default.err ) 2> gpurun_out/r2f_default.time
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_hp_gather -c 2 -o gpurun_out/r2_prof_hpgather2 -f $B --workload c4s --steps 1 --warmup 0 --no-e2e > gpurun_out/r2f_ncu_hpg.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_scatter3d -c 1 -o gpurun_out/r2_prof_scatter3d -f $B --workload c3s --steps 1 --warmup 0 --no-e2e > gpurun_out/r2f_ncu_s3d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_gather2d -c 1 -o gpurun_out/r2_prof_gather2d -f $B --workload c2 --steps 1 --warmup 0 --no-e2e > gpurun_out/r2f_ncu_g2d.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2f_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        def show(d, tag):
            e=d.get("e2e") or {}
            print(tag, "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], d["roofline"]["bound"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()}, "e2e", e.get("ms_per_step"), "parity", (d.get("parity") or {}).get("max_rel_err"), (d.get("parity") or {}).get("ok"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
        show(d, f.split("/")[-1])
        for x in d.get("extra", []):
            if "error" in x: print("   extra ERROR", x)
            else: show(x, "   extra "+x["config"]["name"])
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-600:])
PY
cat gpurun_out/r2f_default.time; ls -la gpurun_out/*.ncu-rep
