#!/bin/bash
# A/B of the launcher's environment knobs on the final code (graph replay, 60 steps each)
mkdir -p gpurun_out
quick() {
  env $1 timeout 600 python bench.py --steps 60 --warmup 5 --skip-cpu-baseline --skip-gpu-baseline --skip-legs 2>gpurun_out/knobs.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$1', 'ms/step', round(d['ms_per_step'],4))"
}
quick "HCA_NOP=1"
quick "HCA_TC_PLDIRECT=1"
quick "HCA_TC_EG=2"
quick "HCA_TC_EG=1"
quick "HCA_TC_PAIR_WGRAD=0"
quick "HCA_TC_PAIR=0"
quick "HCA_SIDE_STREAM=0"
quick "HCA_NOP=2"
