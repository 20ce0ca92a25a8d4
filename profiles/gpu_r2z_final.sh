#!/bin/bash
# Final GPU visit of round 2: full GPU suite, the bench lines (own arm with every leg, reference arm, strong-scaling base), hv ncu capture.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2z_tests.log
timeout 900 python bench.py --steps 200 --warmup 5 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/r2z_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2z_bench_reference_arm.json 2> gpurun_out/r2z_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2z_bench_reference_arm.json
timeout 600 python bench.py --scaling strong --steps 30 --warmup 3 --skip-legs --skip-cpu-baseline --skip-gpu-baseline > gpurun_out/r2z_bench_strong_1gpu.json 2> gpurun_out/r2z_strong.err; echo "strong rc=$?"; cut -c1-200 gpurun_out/r2z_bench_strong_1gpu.json
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:hv_kernel -s 2 -c 2 -o gpurun_out/r2z_hv_full -f python profiles/prof_coattn.py 2 > gpurun_out/r2z_hv_ncu.log 2>&1; echo "hv ncu rc=$?"
ncu -i gpurun_out/r2z_hv_full.ncu-rep --page raw --csv > gpurun_out/r2z_hv_raw.csv 2>/dev/null
python profiles/infer_sweep.py > gpurun_out/r2z_infer_sweep.md 2> gpurun_out/r2z_infer.err; echo "sweep rc=$?"; tail -3 gpurun_out/r2z_infer_sweep.md
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "small_golden or fused_cross_entropy or flat_gradient_sink" > gpurun_out/r2z_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r2z_sanitizer_memcheck.log | head -10
