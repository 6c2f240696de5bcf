#!/bin/bash
# round 2, GPU call J (2 GPUs): multi-GPU tests, torchrun bench at N = 2, in-process multi-GPU render()
mkdir -p gpurun_out
T=$1
(timeout 900 python -m pytest tests -m gpu -q -k "multi_gpu or tone or film_stays or error_conv" 2>&1 | tail -30) > gpurun_out/${T}_tests.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/${T}_bench_n2.log 2> gpurun_out/${T}_bench_n2.err
timeout 600 python tools/inproc_multi_gpu.py --devices 1,2 > gpurun_out/${T}_inproc.log 2> gpurun_out/${T}_inproc.err
QZ_BUILD_TRACE=1 timeout 300 python tools/profile_step.py --workload obj_viewer --spp 4 > gpurun_out/${T}_buildtrace.log 2>&1
tail -5 gpurun_out/${T}_tests.log; cut -c1-400 gpurun_out/${T}_bench_n2.log; tail -3 gpurun_out/${T}_bench_n2.err; cat gpurun_out/${T}_inproc.log; tail -3 gpurun_out/${T}_inproc.err; grep "qz build" gpurun_out/${T}_buildtrace.log
