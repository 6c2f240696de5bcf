#!/bin/bash
# round 2, GPU call AJ: paths deeper than 255 bounces (16-bit depth field, sampler dimension wrap through the late queue)
mkdir -p gpurun_out
(timeout 28 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "deeper" 2>&1 | grep -v "^Loading\|^Render time" | tail -12) > gpurun_out/$1_tests.log 2>&1
cat gpurun_out/$1_tests.log
