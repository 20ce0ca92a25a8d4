#!/bin/bash
# Quick GPU visit: parity tests, warm co-attention timing, the bench line, ncu launch list of the co-attention kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python profiles/prof_coattn.py 5 > gpurun_out/coattn_warm.log 2>&1; tail -3 gpurun_out/coattn_warm.log
timeout 600 python bench.py --steps 20 --warmup 3 --skip-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json | cut -c1-400; tail -3 gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/coattn_launches.csv python profiles/prof_coattn.py 3 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/coattn_launches.csv') if not l.startswith('=='))]
rows=[r for r in rows if r.get('Metric Name')=='gpu__time_duration.sum']
n=len(rows)//3
tot=0
for r in rows[2*n:]:
    v=float(r['Metric Value'].replace(',','')); us=v/1000 if r['Metric Unit'].startswith('n') else v
    tot+=us
    print(f"{us:8.1f} {r['Grid Size']:>12} {r['Kernel Name'][:90]}")
print('total',tot)
PY
