#!/bin/bash
# round 2, GPU call 16: final validation on one GPU — smoke, whole GPU suite, default bench line, launch list of the headline
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2p_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 4 gpurun_out/r2p_smoke.log
timeout 1500 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/r2p_all_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2p_all_tests.log; tail -n 4 gpurun_out/r2p_all_tests.log
( time timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2p_default.json 2> gpurun_out/r2p_default.err ) 2> gpurun_out/r2p_default.time
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2p_reference.json 2> gpurun_out/r2p_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2p_launches_c2.csv python bench.py --steps 2 --warmup 1 --extra none --no-parity --no-cpu-baseline --no-e2e > gpurun_out/r2p_launches_c2.log 2>&1
timeout 600 python bench.py --workload tiny --steps 3 --warmup 2 --extra none --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2p_tiny.json 2> gpurun_out/r2p_tiny.err
timeout 600 python bench.py --accum f32 --steps 3 --warmup 2 --extra none --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2p_c2_f32.json 2> gpurun_out/r2p_c2_f32.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2p_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if d.get("impl")=="reference": print(f.split("/")[-1], d["value"], d["cpu_baseline"]); continue
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
cat gpurun_out/r2p_default.time
