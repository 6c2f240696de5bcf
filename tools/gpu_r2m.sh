#!/bin/bash
# round 2, GPU call M: variants -- no side-line prefetch (cornell), traversal launch bounds / refill threshold (obj, mandelbrot)
mkdir -p gpurun_out
T=$1
B="timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
echo "# cornell default" >> gpurun_out/${T}_bench.log; $B >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# cornell nopf" >> gpurun_out/${T}_bench.log; QZ_LIB_DIR=quetzalcoatlus_b200/_lib_nopf $B >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# opposing nopf" >> gpurun_out/${T}_bench.log; QZ_LIB_DIR=quetzalcoatlus_b200/_lib_nopf $B --workload opposing_planes --spp 32 >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
for V in nopf t6 r4 r16; do
echo "# obj $V" >> gpurun_out/${T}_bench.log; QZ_LIB_DIR=quetzalcoatlus_b200/_lib_$V $B --workload obj_viewer --spp 96 >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
echo "# mandelbrot $V" >> gpurun_out/${T}_bench.log; QZ_LIB_DIR=quetzalcoatlus_b200/_lib_$V $B --workload mandelbrot >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
done
cat gpurun_out/${T}_bench.log | cut -c1-160
