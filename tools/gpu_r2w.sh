#!/bin/bash
# round 2, GPU call W: lazy hit finalisation in the flat closest-hit loop; tests, default bench, flat workloads
mkdir -p gpurun_out
T=$1
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_tests.log 2>&1
(time timeout 600 python bench.py) > gpurun_out/${T}_bench_default.log 2> gpurun_out/${T}_bench_default.err
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline"
r() { echo "# $1" >> gpurun_out/${T}_bench.log; shift; env "$@" >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err; }
r "cornell again" $B
r "glass_spheres 128 spp" $B --workload glass_spheres --spp 128
r "opposing_planes 32 spp" $B --workload opposing_planes --spp 32
r "textures" $B --workload textures
r "mandelbrot" $B --workload mandelbrot
tail -3 gpurun_out/${T}_tests.log; cut -c1-200 gpurun_out/${T}_bench_default.log
