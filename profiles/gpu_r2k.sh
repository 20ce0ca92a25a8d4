#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_gemm.py tests/test_gpu_headline.py -q -m gpu -x --deselect tests/test_gpu_headline.py::test_argmax_agreement_on_10240_samples 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/r2k_tests.log
quick() {
  echo "== $1"
  env $1 timeout 600 python bench.py --steps 40 --warmup 3 --skip-cpu-baseline --skip-gpu-baseline --skip-legs 2>gpurun_out/r2k_bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('ms/step', round(d['ms_per_step'],4))
for r in d['kernel_shares']['top'][:40]:
    if 'gemm_tc_kernel<32' in r['kernel'] or 'hv_kernel' in r['kernel']: print(f\"  {r['us_per_step']:8.1f} {r['launches_per_step']:4.0f}  {r['kernel'][:80]}\")"
}
quick "HCA_NOP=1"
quick "HCA_TC_CK=0"
tail -3 gpurun_out/r2k_bench.err
cp visual-question-answering_b200/libhiecoattn_b200.so /tmp/lib_release.so
HCA_BUILD_TIMELINE=1 python visual-question-answering_b200/build.py --force > /dev/null 2>gpurun_out/r2l_build.err; echo "build rc=$?"
python profiles/timeline_mlp.py 2>&1 | grep -v "CTA 0 k-blocks" | tee gpurun_out/r2l_timeline_mlp_ck.txt
cp /tmp/lib_release.so visual-question-answering_b200/libhiecoattn_b200.so
