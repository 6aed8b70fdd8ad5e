#!/bin/bash
# usage (under gpurun, one GPU): FP32-mode tests + diagnostics, then the whole GPU suite, then the F32 bench line
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_fp32_accumulate.py tests/test_device_group.py -q -m gpu > gpurun_out/gpu3_new.log 2>&1
echo "pytest exit $?" >> gpurun_out/gpu3_new.log; tail -4 gpurun_out/gpu3_new.log
timeout 100 python tools/fp32_diag.py > gpurun_out/fp32_diag3.jsonl 2> gpurun_out/fp32_diag3.err
python - <<'PY'
import json
for l in open('gpurun_out/fp32_diag3.jsonl'):
    d = json.loads(l)
    print(d['kernel'], 'Q %.3f' % d['quantity']['worst_bar'], d['quantity']['pixel'], 'W %.3f' % d['weight']['worst_bar'],
          d['weight']['pixel'], d['weight']['rel_q50_q99_max(px > 1e-3 max)'])
PY
timeout 400 python -m pytest tests -q -m gpu > gpurun_out/gpu3_all.log 2>&1
echo "pytest exit $?" >> gpurun_out/gpu3_all.log; tail -4 gpurun_out/gpu3_all.log
timeout 150 python bench.py --steps 2 --warmup 3 --accum f32 --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_f32_v2.json 2> gpurun_out/bench_c2_f32_v2.err
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_f32_v2.json')); print('f32 mode C2:', d['value'], d['ms_per_step'], d['roofline']['fp64']['phase_ms'])"
