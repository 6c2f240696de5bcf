#!/bin/bash
# round 2, GPU call Q (2 GPUs): class-compacted memo -- tests, N = 1 / 2 bench, traversal variants
mkdir -p gpurun_out
T=$1
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_tests.log 2>&1
B="timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-secondary"
r() { echo "# $1" >> gpurun_out/${T}_bench.log; shift; env "$@" >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err; }
r "cornell N1" $B
echo "# cornell N2" >> gpurun_out/${T}_bench.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err
r "obj default" $B --workload obj_viewer --spp 96
r "obj leaf4" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_leaf4 $B --workload obj_viewer --spp 96 --pool 16777216
r "obj push" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_push $B --workload obj_viewer --spp 96 --pool 16777216
r "mandelbrot default" $B --workload mandelbrot
r "mandelbrot leaf4" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_leaf4 $B --workload mandelbrot --pool 16777216
r "mandelbrot push" QZ_LIB_DIR=quetzalcoatlus_b200/_lib_push $B --workload mandelbrot --pool 16777216
timeout 300 python tools/inproc_multi_gpu.py --devices 1,2 > gpurun_out/${T}_inproc.jsonl 2>> gpurun_out/${T}_bench.err
tail -4 gpurun_out/${T}_tests.log; grep -v "^$" gpurun_out/${T}_bench.log | cut -c1-160; cut -c1-200 gpurun_out/${T}_inproc.jsonl
