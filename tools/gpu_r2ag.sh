#!/bin/bash
# round 2, GPU call AG (last): final tree -- GPU tests, smoke, default bench, launch list and counters of the headline workload
mkdir -p gpurun_out
T=$1
(timeout 150 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/${T}_tests.log 2>&1
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1
(time timeout 120 python bench.py) > gpurun_out/${T}_bench_default.log 2> gpurun_out/${T}_bench_default.err
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_cornell_box.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/${T}_launches.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
QZ_GRAPH=0 QZ_PIPELINES=1 timeout 70 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_counters_cornell_box.csv \
    python tools/profile_step.py --workload cornell_box --spp 12 > gpurun_out/${T}_counters_cornell_box.log 2>&1
tail -3 gpurun_out/${T}_tests.log; cat gpurun_out/${T}_smoke.log | tail -1; cut -c1-200 gpurun_out/${T}_bench_default.log | tail -2; tail -4 gpurun_out/${T}_bench_default.err
