#!/bin/bash
# checkpoint visit: full GPU suite + a short bench with the per-kernel breakdown
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_headline.py::test_argmax_agreement_on_10240_samples 2>&1 | tail -4 | cut -c1-300 | tee gpurun_out/r2m_tests.log
timeout 600 python bench.py --steps 60 --warmup 5 --skip-cpu-baseline --skip-gpu-baseline --skip-legs 2>gpurun_out/r2m_bench.err | tee gpurun_out/r2m_bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],4), 'samples/s', round(d['value']), 'launches/step', d.get('gpu_launches_per_step'))
for r in d['kernel_shares']['top'][:45]: print(f\"  {r['us_per_step']:8.1f} {r['launches_per_step']:4.0f}  {r['kernel'][:90]}\")"
