#!/bin/bash
# round 2, GPU call AF (8 GPUs): BASELINE config 5 -- cornell 3840x2160, 1024 spp over 8 GPUs -- with the final code
mkdir -p gpurun_out
T=$1
timeout 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29881 bench.py --gpus 8 --workload cornell_4k --steps 2 --warmup 1 --no-cpu-baseline --no-secondary >> gpurun_out/${T}_scale_cornell_4k.jsonl 2>> gpurun_out/${T}_scale.err
cut -c1-140 gpurun_out/${T}_scale_cornell_4k.jsonl; tail -2 gpurun_out/${T}_scale.err
