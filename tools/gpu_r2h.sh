#!/bin/bash
# round 2, GPU call H: the new bench.py end to end; steady-state captures of the traversal kernels on the BVH workloads
mkdir -p gpurun_out
T=$1
(time timeout 600 python bench.py) > gpurun_out/${T}_bench_default.log 2> gpurun_out/${T}_bench_default.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.log 2>> gpurun_out/${T}_bench_default.err
for W in obj_viewer mandelbrot; do
QZ_GRAPH=0 QZ_PIPELINES=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_trace_lane --launch-skip 8 --launch-count 2 \
    -o gpurun_out/${T}_${W}_trace -f python tools/profile_step.py --workload $W --spp 48 > gpurun_out/${T}_ncu_${W}.log 2>&1
done
cut -c1-600 gpurun_out/${T}_bench_default.log; tail -5 gpurun_out/${T}_bench_default.err; cut -c1-300 gpurun_out/${T}_bench_reference.log
