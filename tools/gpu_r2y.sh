#!/bin/bash
# round 2, GPU call Y (4 GPUs): N = 4 after moving the clock sampler to rank 0 only
mkdir -p gpurun_out
T=$1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29840 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary >> gpurun_out/${T}_scale_cornell_box.jsonl 2>> gpurun_out/${T}_scale.err
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary >> gpurun_out/${T}_scale_cornell_box.jsonl 2>> gpurun_out/${T}_scale.err
cut -c1-140 gpurun_out/${T}_scale_cornell_box.jsonl; tail -2 gpurun_out/${T}_scale.err
