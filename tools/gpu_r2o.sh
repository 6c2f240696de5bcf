#!/bin/bash
# round 2, GPU call O: prefetch variants of k_shade / k_step_flat
mkdir -p gpurun_out
T=$1
B="timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
r() { echo "# $1" >> gpurun_out/${T}_bench.log; shift; env "$@" >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err; }
r "cornell hot" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_hot $B
r "cornell hotl1" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_hotl1 $B
r "cornell hotl1 P2" QZ_PIPELINES=2 QZ_LIB_DIR=quetzalcoatlus_b200/_lib_hotl1 $B
r "cornell p1" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_p1 $B
r "opposing hotl1" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_hotl1 $B --workload opposing_planes --spp 32
r "glass hotl1" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_hotl1 $B --workload glass_spheres --spp 128
r "glass p1" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_p1 $B --workload glass_spheres --spp 128
r "obj hotl1 P2 pool24" QZ_PIPELINES=2 QZ_LIB_DIR=quetzalcoatlus_b200/_lib_hotl1 $B --workload obj_viewer --spp 96 --pool 16777216
r "mandelbrot P2" QZ_PIPELINES=2 $B --workload mandelbrot
cat gpurun_out/${T}_bench.log | cut -c1-160
