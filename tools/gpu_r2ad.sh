#!/bin/bash
# round 2, GPU call AD: the default bench line with the final bench.py (mesh_first_hit_fraction in the secondary obj_viewer line)
mkdir -p gpurun_out
T=$1
(time timeout 600 python bench.py) > gpurun_out/${T}_bench_default.log 2> gpurun_out/${T}_bench_default.err
timeout 300 python bench.py --workload obj_viewer --steps 3 --warmup 3 --no-secondary > gpurun_out/${T}_bench_obj.log 2>> gpurun_out/${T}_bench_default.err
cut -c1-200 gpurun_out/${T}_bench_default.log; tail -5 gpurun_out/${T}_bench_default.err; cut -c1-200 gpurun_out/${T}_bench_obj.log
