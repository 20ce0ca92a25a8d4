#!/bin/bash
# GPU visit r2d (8 GPUs): data-parallel bench -- fused NVLink kernel (with / without the early slice), NCCL transport, strong scaling at N = 8.
mkdir -p gpurun_out
run() {  # name, env, extra args
  env $2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) \
    bench.py --gpus 8 --steps 100 --warmup 5 --skip-legs $3 > gpurun_out/r2d_$1.json 2> gpurun_out/r2d_$1.err
  echo "$1 rc=$?"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2d_$1.json'))
    print('  ', '$1', 'ms/step', round(d['ms_per_step'],4), 'samples/s', round(d['value']), 'e2e', round(d['e2e']['value']), d['config']['collective'][:90])
except Exception as e:
    print('   no json:', e)
PY
}
run fused "HCA_X=1" ""
run fused_noearly "HCA_X=1" "--no-early-reduce"
run nccl "HCA_DP_FUSED=0" ""
run strong "HCA_X=1" "--scaling strong"
tail -3 gpurun_out/r2d_fused.err
