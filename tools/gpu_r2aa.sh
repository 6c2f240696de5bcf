#!/bin/bash
# round 2, GPU call AA: last sweeps (conductor / run-time-dispatch shade kernels at 3 CTAs per SM; pipelines on the BVH workloads; lean CTAs)
mkdir -p gpurun_out
T=$1
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline"
r() { echo "# $1" >> gpurun_out/${T}_bench.log; shift; env "$@" >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err; }
r "opposing default" $B --workload opposing_planes --spp 32
r "opposing heavy3" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_heavy3 $B --workload opposing_planes --spp 32
r "cornell heavy3" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_heavy3 $B
r "cornell lean2" QZ_LEAN_BLOCKS_PER_SM=2 $B
r "cornell default" $B
r "cornell lean2 again" QZ_LEAN_BLOCKS_PER_SM=2 $B
r "obj default" $B --workload obj_viewer --spp 96
r "obj P2" QZ_PIPELINES=2 $B --workload obj_viewer --spp 96
r "obj lean2" QZ_LEAN_BLOCKS_PER_SM=2 $B --workload obj_viewer --spp 96
r "mandelbrot P2" QZ_PIPELINES=2 $B --workload mandelbrot
cat gpurun_out/${T}_bench.log | cut -c1-140
