#!/bin/bash
# usage (under gpurun, one GPU): parity of the 2D paths after the k_gather2d changes (tile-relative records, scaled
# polynomial, packed-FP32 chain), the C2 bench lines in FP64 and FP32-accumulate mode, one ncu capture of the
# FP32-accumulate instantiation of k_gather2d
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity_2d.py tests/test_fp32_accumulate.py tests/test_golden_vectors.py \
    tests/test_gpu_sedov_config.py -q -m gpu > gpurun_out/gpu4_tests.log 2>&1
echo "pytest exit $?" >> gpurun_out/gpu4_tests.log; tail -8 gpurun_out/gpu4_tests.log
timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_v4.json 2> gpurun_out/bench_c2_v4.err
timeout 100 python bench.py --steps 3 --warmup 3 --accum f32 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_f32_v4.json 2> gpurun_out/bench_c2_f32_v4.err
python - <<'PY'
import json
for f in ("bench_c2_v4", "bench_c2_f32_v4"):
    try:
        d = json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d["value"], 3), "Mp/s", round(d["ms_per_step"], 1), "ms",
              {k: round(v, 1) for k, v in d["roofline"]["fp64"]["phase_ms"].items()}, d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
if [ "$1" == "ncu" ]; then
timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_gather2d -c 1 -o gpurun_out/prof_g2d_f32 -f \
    python bench.py --workload small --steps 1 --warmup 0 --accum f32 --no-e2e --no-cpu-baseline > gpurun_out/ncu_g2d_f32.log 2>&1
ls -la gpurun_out/prof_g2d_f32.ncu-rep
fi
