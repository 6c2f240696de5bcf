#!/bin/bash
# round 2, GPU call P (8 GPUs): scaling lines -- torchrun N = 1, 2, 4, 8 (default weak config and cornell_4k), in-process multi-GPU render()
mkdir -p gpurun_out
T=$1
run() { n=$1; shift; if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n "$@"; fi; }
for n in 1 2 4 8; do run $n --steps 5 --warmup 3 --no-cpu-baseline --no-secondary >> gpurun_out/${T}_scale_cornell_box.jsonl 2>> gpurun_out/${T}_scale.err; done
for n in 1 8; do run $n --workload cornell_4k --steps 3 --warmup 1 --no-cpu-baseline --no-secondary >> gpurun_out/${T}_scale_cornell_4k.jsonl 2>> gpurun_out/${T}_scale.err; done
timeout 600 python tools/inproc_multi_gpu.py --devices 1,2,4,8 > gpurun_out/${T}_inproc.jsonl 2>> gpurun_out/${T}_scale.err
timeout 300 python -m pytest tests -m gpu -q -k "multi_gpu" 2>&1 | tail -3 > gpurun_out/${T}_tests.log
grep -c . gpurun_out/${T}_scale_cornell_box.jsonl; cut -c1-120 gpurun_out/${T}_scale_cornell_box.jsonl; cut -c1-120 gpurun_out/${T}_scale_cornell_4k.jsonl; cut -c1-200 gpurun_out/${T}_inproc.jsonl; cat gpurun_out/${T}_tests.log; tail -5 gpurun_out/${T}_scale.err
