#!/bin/bash
# round 2, GPU call X: small sweeps on the final kernels (flat tile size, barrier-free memo fill, pipelines, CTAs per SM)
mkdir -p gpurun_out
T=$1
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline"
r() { echo "# $1" >> gpurun_out/${T}_bench.log; shift; env "$@" >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err; }
r "cornell default" $B
r "cornell tile512" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_tile512 $B
r "cornell tile2048" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_tile2048 $B
r "cornell fill direct" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_filld $B
r "cornell P2" QZ_PIPELINES=2 $B
r "cornell lean 4" QZ_LEAN_BLOCKS_PER_SM=4 $B
r "cornell lean 2" QZ_LEAN_BLOCKS_PER_SM=2 $B
r "cornell blocks 4" QZ_BLOCKS_PER_SM=4 $B
cat gpurun_out/${T}_bench.log | cut -c1-140
