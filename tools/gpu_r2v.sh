#!/bin/bash
# round 2, GPU call V: lean no-material shade kernel; tests, default bench, BVH workload lines
mkdir -p gpurun_out
T=$1
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_tests.log 2>&1
(time timeout 600 python bench.py) > gpurun_out/${T}_bench_default.log 2> gpurun_out/${T}_bench_default.err
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline"
r() { echo "# $1" >> gpurun_out/${T}_bench.log; shift; env "$@" >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err; }
r "cornell again" $B
r "obj_viewer 96 spp" $B --workload obj_viewer --spp 96
r "mandelbrot" $B --workload mandelbrot
r "opposing_planes 32 spp" $B --workload opposing_planes --spp 32
r "textures" $B --workload textures
tail -3 gpurun_out/${T}_tests.log; cut -c1-200 gpurun_out/${T}_bench_default.log
