#!/bin/bash
# round 2, GPU call A: tests, bench lines (default / exact / secondary workloads), one ncu capture of a cornell_box iteration
mkdir -p gpurun_out
T=$1
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40) > gpurun_out/${T}_tests.log 2>&1
B="timeout 400 python bench.py --steps 5 --warmup 3"
echo "# default" >> gpurun_out/${T}_bench.log; $B >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# exact" >> gpurun_out/${T}_bench.log; $B --no-cpu-baseline --flags 64 >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# nograph" >> gpurun_out/${T}_bench.log; QZ_GRAPH=0 $B --no-cpu-baseline >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# obj" >> gpurun_out/${T}_bench.log; $B --no-cpu-baseline --workload obj_viewer --spp 96 >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# mandelbrot" >> gpurun_out/${T}_bench.log; $B --no-cpu-baseline --workload mandelbrot >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# glass" >> gpurun_out/${T}_bench.log; $B --no-cpu-baseline --workload glass_spheres --spp 128 >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# opposing" >> gpurun_out/${T}_bench.log; $B --no-cpu-baseline --workload opposing_planes --spp 32 >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
# one steady-state iteration of one pipeline, kernel by kernel
QZ_GRAPH=0 QZ_PIPELINES=1 timeout 600 ncu --set full --clock-control none --import-source on --launch-skip 81 --launch-count 16 \
    -o gpurun_out/${T}_cornell_iter -f python tools/profile_step.py --workload cornell_box --spp 24 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_bench.log | cut -c1-300
