#!/bin/bash
# round 2, GPU call S: regression after the strip rule / class-table cache; shard probe; default bench
mkdir -p gpurun_out
T=$1
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_tests.log 2>&1
timeout 100 python tools/shard_probe.py > gpurun_out/${T}_shard_probe.jsonl 2>> gpurun_out/${T}_bench.err
timeout 100 python tools/shard_probe.py --workload cornell_4k --shards 8 --steps 2 >> gpurun_out/${T}_shard_probe.jsonl 2>> gpurun_out/${T}_bench.err
(time timeout 600 python bench.py) > gpurun_out/${T}_bench_default.log 2> gpurun_out/${T}_bench_default.err
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu-baseline"
r() { echo "# $1" >> gpurun_out/${T}_bench.log; shift; env "$@" >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err; }
r "textures" $B --workload textures
r "mandelbrot" $B --workload mandelbrot
r "glass bvh" $B --workload glass_spheres --spp 128 --flags 8
r "glass flat" $B --workload glass_spheres --spp 128
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_shard_probe.jsonl; cut -c1-200 gpurun_out/${T}_bench_default.log
