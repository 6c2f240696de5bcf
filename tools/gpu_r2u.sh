#!/bin/bash
# round 2, GPU call U (4 GPUs): scaling with the in-process NVML clock sampler
mkdir -p gpurun_out
T=$1
run() { n=$1; shift; if [ $n = 1 ]; then timeout 300 python bench.py --gpus 1 "$@"; else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 297$n bench.py --gpus $n "$@"; fi; }
for n in 1 2 4; do run $n --steps 5 --warmup 3 --no-cpu-baseline --no-secondary >> gpurun_out/${T}_scale_cornell_box.jsonl 2>> gpurun_out/${T}_scale.err; done
run 4 --workload cornell_4k --steps 2 --warmup 1 --no-cpu-baseline --no-secondary >> gpurun_out/${T}_scale_cornell_4k.jsonl 2>> gpurun_out/${T}_scale.err
cut -c1-120 gpurun_out/${T}_scale_cornell_box.jsonl; cut -c1-120 gpurun_out/${T}_scale_cornell_4k.jsonl; tail -3 gpurun_out/${T}_scale.err
