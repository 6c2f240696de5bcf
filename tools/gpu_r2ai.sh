#!/bin/bash
# round 2, GPU call AI: the tone kernel after its 8-bit conversion was pinned against OpenCV (NaN / inf / beyond int32 -> 0)
mkdir -p gpurun_out
(timeout 50 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "tone or film_stays" 2>&1 | tail -6) > gpurun_out/$1_tests.log 2>&1
cat gpurun_out/$1_tests.log
