#!/bin/bash
# round 2, GPU call K: late queue, slab tweak, lazy pop, copy-out; tests + full bench + per-workload lines
mkdir -p gpurun_out
T=$1
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_tests.log 2>&1
(time QZ_BUILD_TRACE=1 timeout 600 python bench.py) > gpurun_out/${T}_bench_default.log 2> gpurun_out/${T}_bench_default.err
B="timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
echo "# obj" >> gpurun_out/${T}_bench.log; $B --workload obj_viewer --spp 96 >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# mandelbrot" >> gpurun_out/${T}_bench.log; $B --workload mandelbrot >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# glass" >> gpurun_out/${T}_bench.log; $B --workload glass_spheres --spp 128 >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# opposing" >> gpurun_out/${T}_bench.log; $B --workload opposing_planes --spp 32 >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# textures" >> gpurun_out/${T}_bench.log; $B --workload textures >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
QZ_GRAPH=0 QZ_PIPELINES=1 timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_step_flat|k_shade|k_sample|k_bin|k_albedo" --launch-skip 32 --launch-count 8 \
    -o gpurun_out/${T}_cornell_steady -f python tools/profile_step.py --workload cornell_box --spp 64 > gpurun_out/${T}_ncu.log 2>&1
tail -3 gpurun_out/${T}_tests.log; cut -c1-300 gpurun_out/${T}_bench_default.log; cat gpurun_out/${T}_bench.log | cut -c1-200
