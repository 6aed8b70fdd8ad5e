#!/bin/bash
# round 2, GPU call 7 (gpurun --gpus 2): N>1 correctness (tools/mgpu_check.py) and the default bench line under torchrun
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $T tools/mgpu_check.py > gpurun_out/r2g_mgpu_check.log 2>&1; echo "mgpu_check rc=$?"; tail -n 8 gpurun_out/r2g_mgpu_check.log
( time timeout 1500 $T bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2g_default_2gpu.json 2> gpurun_out/r2g_default_2gpu.err ) 2> gpurun_out/r2g_default_2gpu.time
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r2g_*.json")):
    try:
        d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
        def show(d, tag):
            e=d.get("e2e") or {}
            print(tag, "N=%d"%d["n_gpus"], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], d["roofline"]["bound"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()}, "e2e", e.get("ms_per_step"), d["config"].get("shard"))
        show(d, f.split("/")[-1])
        for x in d.get("extra", []):
            if "error" in x: print("   extra ERROR", x)
            else: show(x, "   extra "+x["config"]["name"])
    except Exception as ex:
        print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-1500:])
PY
cat gpurun_out/r2g_default_2gpu.time
