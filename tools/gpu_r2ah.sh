#!/bin/bash
# round 2, GPU call AH: last sanity of the committed tree -- smoke and a short device-timed bench
mkdir -p gpurun_out
T=$1
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
timeout 45 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${T}_bench.log 2> gpurun_out/${T}_bench.err
tail -1 gpurun_out/${T}_smoke.log; cut -c1-160 gpurun_out/${T}_bench.log; tail -2 gpurun_out/${T}_bench.err
