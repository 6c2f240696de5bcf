#!/bin/bash
# round 2, GPU call AC: light emission preloaded at the top of the bounce (variant _lib_pre)
mkdir -p gpurun_out
T=$1
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline"
r() { echo "# $1" >> gpurun_out/${T}_bench.log; shift; env "$@" >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err; }
r "cornell default" $B
r "cornell pre" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_pre $B
r "opposing pre" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_pre $B --workload opposing_planes --spp 32
r "cornell pre again" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_pre $B
cat gpurun_out/${T}_bench.log | cut -c1-140
