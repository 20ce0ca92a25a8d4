#!/bin/bash
# quick GPU check: GEMM + parity tests, then the bench line (no baselines)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_gpu_lstm.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --steps 100 --warmup 5 --skip-cpu-baseline --skip-gpu-baseline > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/quick_bench.json'))
print('ms/step', round(d['ms_per_step'],4), 'samples/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches_per_step'])
for l in d['roofline_legs']: print('  leg', l.get('name'), round(l.get('ms_per_launch',0)*1e3,1),'us frac', round(l.get('frac',0),4), 'issued', round(l.get('frac_issued',0),3))
for r in d['kernel_shares']['top'][:14]: print(f"  {r['us_per_step']:8.1f} {r['launches_per_step']:4.0f}  {r['kernel'][:80]}")
PY
