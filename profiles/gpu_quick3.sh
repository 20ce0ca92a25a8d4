#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x -s --deselect tests/test_gpu_headline.py::test_argmax_agreement_on_10240_samples 2>&1 | grep -E "B=160|pool indices|passed|failed|Error|worst|FlatAdam" | cut -c1-220
timeout 600 python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --skip-gpu-baseline --skip-legs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],4), 'samples/s', round(d['value']))
for r in d['kernel_shares']['top'][:12]: print(f\"  {r['us_per_step']:8.1f} {r['launches_per_step']:4.0f}  {r['kernel'][:80]}\")"
