#!/bin/bash
# round 2, GPU call I: cheaper slab test; memo-row prefetch variant
mkdir -p gpurun_out
T=$1
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_tests.log 2>&1
B="timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
echo "# default" >> gpurun_out/${T}_bench.log; $B >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# lib_pf" >> gpurun_out/${T}_bench.log; QZ_LIB_DIR=quetzalcoatlus_b200/_lib_pf $B >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# obj" >> gpurun_out/${T}_bench.log; $B --workload obj_viewer --spp 96 >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# mandelbrot" >> gpurun_out/${T}_bench.log; $B --workload mandelbrot >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# opposing pf" >> gpurun_out/${T}_bench.log; QZ_LIB_DIR=quetzalcoatlus_b200/_lib_pf $B --workload opposing_planes --spp 32 >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
W=obj_viewer
QZ_GRAPH=0 QZ_PIPELINES=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:k_trace_lane --launch-skip 8 --launch-count 2 \
    -o gpurun_out/${T}_${W}_trace -f python tools/profile_step.py --workload $W --spp 48 > gpurun_out/${T}_ncu_${W}.log 2>&1
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_bench.log | cut -c1-200
