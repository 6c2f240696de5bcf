mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r1d_tests.log 2>&1
B="timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
echo "# default" >> gpurun_out/r1d_bench.log; $B >> gpurun_out/r1d_bench.log 2>>gpurun_out/r1d_bench.err
echo "# lib_a (6/4)" >> gpurun_out/r1d_bench.log; QZ_LIB_DIR=quetzalcoatlus_b200/_lib_a $B >> gpurun_out/r1d_bench.log 2>>gpurun_out/r1d_bench.err
echo "# lib_b (4/3)" >> gpurun_out/r1d_bench.log; QZ_LIB_DIR=quetzalcoatlus_b200/_lib_b $B >> gpurun_out/r1d_bench.log 2>>gpurun_out/r1d_bench.err
echo "# default force-bvh" >> gpurun_out/r1d_bench.log; $B --flags 8 >> gpurun_out/r1d_bench.log 2>>gpurun_out/r1d_bench.err
echo "# default unsorted" >> gpurun_out/r1d_bench.log; $B --flags 1 >> gpurun_out/r1d_bench.log 2>>gpurun_out/r1d_bench.err
echo "# glass flat" >> gpurun_out/r1d_bench.log; $B --workload glass_spheres --spp 128 >> gpurun_out/r1d_bench.log 2>>gpurun_out/r1d_bench.err
echo "# glass bvh" >> gpurun_out/r1d_bench.log; $B --workload glass_spheres --spp 128 --flags 8 >> gpurun_out/r1d_bench.log 2>>gpurun_out/r1d_bench.err
echo "# opposing" >> gpurun_out/r1d_bench.log; $B --workload opposing_planes --spp 32 >> gpurun_out/r1d_bench.log 2>>gpurun_out/r1d_bench.err
echo "# textures" >> gpurun_out/r1d_bench.log; $B --workload textures >> gpurun_out/r1d_bench.log 2>>gpurun_out/r1d_bench.err
echo "# obj 1M" >> gpurun_out/r1d_bench.log; $B --workload obj_viewer >> gpurun_out/r1d_bench.log 2>>gpurun_out/r1d_bench.err
tail -3 gpurun_out/r1d_tests.log
