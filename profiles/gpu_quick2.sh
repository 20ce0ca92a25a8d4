#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_headline.py tests/test_gpu_parity.py -q -m gpu -x --deselect tests/test_gpu_headline.py::test_argmax_agreement_on_10240_samples 2>&1 | tail -3
timeout 300 python profiles/launch_order.py 2>&1 | grep -E "fixup|pool3_fwd|conv_norms|total"
