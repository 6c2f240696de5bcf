#!/bin/bash
# round 2, GPU call AB: CTAs per SM of the stage kernels with three pipelines (lean stages now default to 2)
mkdir -p gpurun_out
T=$1
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline"
r() { echo "# $1" >> gpurun_out/${T}_bench.log; shift; env "$@" >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err; }
r "cornell default (lean 2)" $B
r "cornell blocks 2" QZ_BLOCKS_PER_SM=2 $B
r "cornell lean 1" QZ_LEAN_BLOCKS_PER_SM=1 $B
r "cornell blocks 2 lean 1" QZ_BLOCKS_PER_SM=2 QZ_LEAN_BLOCKS_PER_SM=1 $B
r "textures default" $B --workload textures
r "glass default" $B --workload glass_spheres --spp 128
cat gpurun_out/${T}_bench.log | cut -c1-140
