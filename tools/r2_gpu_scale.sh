#!/bin/bash
# usage (under gpurun --gpus N): tools/r2_gpu_scale.sh N   — the default bench line under torchrun on N GPUs
N=$1
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"
( time timeout 1500 $T bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2q_default_${N}gpu.json 2> gpurun_out/r2q_default_${N}gpu.err ) 2> gpurun_out/r2q_default_${N}gpu.time
python - <<PY
import json
f="gpurun_out/r2q_default_${N}gpu.json"
try:
    d=json.loads([l for l in open(f).read().strip().splitlines() if l.startswith("{")][-1])
    def show(d, tag):
        e=d.get("e2e") or {}
        print(tag, "N=%d"%d["n_gpus"], "%.2f Mp/s %.1f ms"%(d["value"],d["ms_per_step"]), d["roofline"]["kernel"], "frac %.3f"%d["roofline"]["frac"], {k:round(v,1) for k,v in d["roofline"]["phase_ms"].items()}, "e2e", e.get("ms_per_step"))
    show(d, "main")
    for x in d.get("extra", []):
        if "error" in x: print("   extra ERROR", x)
        else: show(x, "   extra "+x["config"]["name"])
except Exception as ex:
    print(f, "ERR", ex, open(f.replace(".json",".err")).read()[-1500:])
PY
cat gpurun_out/r2q_default_${N}gpu.time
