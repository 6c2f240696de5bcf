#!/bin/bash
# round 2, GPU call N: tuning sweeps on the final kernels
mkdir -p gpurun_out
T=$1
B="timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
r() { echo "# $1" >> gpurun_out/${T}_bench.log; shift; env "$@" >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err; }
r "cornell default" $B
r "cornell l1" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_l1 $B
r "opposing l1" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_l1 $B --workload opposing_planes --spp 32
r "cornell pool 2^22" $B --pool 4194304
r "cornell pool 2^24" $B --pool 16777216
r "cornell pipelines 2" QZ_PIPELINES=2 $B
r "cornell pipelines 4" QZ_PIPELINES=4 $B
r "cornell memo 6 bounces" QZ_MEMO_BOUNCES=6 $B
r "cornell memo 12 bounces" QZ_MEMO_BOUNCES=12 $B
r "obj r24" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_r24 $B --workload obj_viewer --spp 96
r "mandelbrot r24" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_r24 $B --workload mandelbrot
r "obj pipelines 2" QZ_PIPELINES=2 $B --workload obj_viewer --spp 96
r "obj pipelines 4" QZ_PIPELINES=4 $B --workload obj_viewer --spp 96
r "obj pool 2^24" $B --workload obj_viewer --spp 96 --pool 16777216
cat gpurun_out/${T}_bench.log | cut -c1-160
