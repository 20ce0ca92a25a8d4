#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_headline.py -q -m gpu -x -s --deselect tests/test_gpu_headline.py::test_argmax_agreement_on_10240_samples 2>&1 | grep -E "B=160|pool indices|passed|failed|Error|error|assert" | cut -c1-220 | tee gpurun_out/r2j_tests.log
quick() {
  echo "== $1"
  env $1 timeout 600 python bench.py --steps 40 --warmup 3 --skip-cpu-baseline --skip-gpu-baseline --skip-legs 2>gpurun_out/r2j_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],4))
for r in d['kernel_shares']['top'][:40]:
    if 'hv_kernel' in r['kernel']: print(f\"  {r['us_per_step']:8.1f} {r['launches_per_step']:4.0f}  {r['kernel'][:80]}\")"
}
quick "HCA_NOP=1"
quick "HCA_HV_BN=128"
cp visual-question-answering_b200/libhiecoattn_b200.so /tmp/lib_release.so
HCA_BUILD_TIMELINE=1 python visual-question-answering_b200/build.py --force > /dev/null 2>gpurun_out/r2j_build.err; echo "build rc=$?"
python profiles/timeline_hv.py 2>&1 | tee gpurun_out/r2j_timeline_bn64.txt
cp /tmp/lib_release.so visual-question-answering_b200/libhiecoattn_b200.so
