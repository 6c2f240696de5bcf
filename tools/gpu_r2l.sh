#!/bin/bash
# round 2, GPU call L: tests, default bench, ncu launch list of the bench command, ncu counters per unit of work
mkdir -p gpurun_out
T=$1
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_tests.log 2>&1
(time timeout 600 python bench.py) > gpurun_out/${T}_bench_default.log 2> gpurun_out/${T}_bench_default.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches_cornell_box.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/${T}_launches.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
for W in cornell_box:12 obj_viewer:8 mandelbrot:8; do
  N=${W%%:*}; S=${W##*:}
  QZ_GRAPH=0 QZ_PIPELINES=1 timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_counters_${N}.csv \
      python tools/profile_step.py --workload $N --spp $S > gpurun_out/${T}_counters_${N}.log 2>&1
done
tail -3 gpurun_out/${T}_tests.log; cut -c1-300 gpurun_out/${T}_bench_default.log; ls -la gpurun_out/${T}_*
