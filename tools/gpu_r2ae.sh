#!/bin/bash
# round 2, GPU call AE (8 GPUs): N = 8 with the final bench.py and library
mkdir -p gpurun_out
T=$1
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29880 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary >> gpurun_out/${T}_scale_cornell_box.jsonl 2>> gpurun_out/${T}_scale.err
cut -c1-140 gpurun_out/${T}_scale_cornell_box.jsonl; tail -2 gpurun_out/${T}_scale.err
