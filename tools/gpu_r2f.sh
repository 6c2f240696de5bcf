#!/bin/bash
# round 2, GPU call F: steady-state ncu captures of the BVH workloads (one iteration of one pipeline each)
mkdir -p gpurun_out
T=$1
QZ_GRAPH=0 QZ_PIPELINES=1 timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 32 --launch-count 10 \
    -o gpurun_out/${T}_obj_steady -f python tools/profile_step.py --workload obj_viewer --spp 48 > gpurun_out/${T}_ncu_obj.log 2>&1
QZ_GRAPH=0 QZ_PIPELINES=1 timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 22 --launch-count 10 \
    -o gpurun_out/${T}_mandel_steady -f python tools/profile_step.py --workload mandelbrot --spp 32 > gpurun_out/${T}_ncu_mandel.log 2>&1
tail -2 gpurun_out/${T}_ncu_obj.log gpurun_out/${T}_ncu_mandel.log
