#!/bin/bash
# round 2, GPU call AK: Image::save("*.png") through the product library (tone kernel + PNG encoder)
mkdir -p gpurun_out
(timeout 12 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "image_save_png" 2>&1 | tail -5) > gpurun_out/$1_tests.log 2>&1
cat gpurun_out/$1_tests.log
