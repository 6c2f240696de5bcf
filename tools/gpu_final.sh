#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, both bench arms
mkdir -p gpurun_out
T=$1
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5) > gpurun_out/${T}_tests.log 2>&1
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/${T}_smoke.log 2>&1
(time timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1) > gpurun_out/${T}_bench_reference.log 2> gpurun_out/${T}_bench_reference.err
(time timeout 600 python bench.py) > gpurun_out/${T}_bench_default.log 2> gpurun_out/${T}_bench_default.err
cat gpurun_out/${T}_tests.log gpurun_out/${T}_smoke.log; cut -c1-160 gpurun_out/${T}_bench_reference.log; cut -c1-160 gpurun_out/${T}_bench_default.log; tail -4 gpurun_out/${T}_bench_default.err
