#!/bin/bash
# Short GPU visit: parity tests, the bench line (no CPU leg), optional extra command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches_per_step','final_loss')}, d['e2e']['value'], d['roofline']['ms_per_launch'], d['config']['mode'])
    for r in d['kernel_shares']['top']: print('  ', round(r['us_per_step'],1), r['launches_per_step'], r['kernel'][:70])
    print('kernel_us_per_step', d['kernel_shares']['kernel_us_per_step'])
except Exception as e: print('bench parse failed', e)
PY
tail -5 gpurun_out/bench.err | cut -c1-300
if [ -n "$1" ]; then timeout 600 bash -c "$1" > gpurun_out/extra.log 2>&1; tail -30 gpurun_out/extra.log; fi
