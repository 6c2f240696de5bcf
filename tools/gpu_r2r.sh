#!/bin/bash
# round 2, GPU call R: final code -- tests, bench (both arms), per-workload lines, launch list, counters, steady-state captures, shard probe
mkdir -p gpurun_out
T=$1
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30) > gpurun_out/${T}_tests.log 2>&1
(time timeout 600 python bench.py) > gpurun_out/${T}_bench_default.log 2> gpurun_out/${T}_bench_default.err
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.log 2>> gpurun_out/${T}_bench_default.err
B="timeout 200 python bench.py --steps 5 --warmup 3 --no-secondary"
r() { echo "# $1" >> gpurun_out/${T}_bench.log; shift; env "$@" >> gpurun_out/${T}_bench.log 2>>gpurun_out/${T}_bench.err; }
r "exact arithmetic" $B --no-cpu-baseline --flags 64
r "obj_viewer 96 spp" $B --workload obj_viewer --spp 96
r "mandelbrot" $B --workload mandelbrot --no-cpu-baseline
r "glass_spheres 128 spp" $B --workload glass_spheres --spp 128 --no-cpu-baseline
r "opposing_planes 32 spp" $B --workload opposing_planes --spp 32 --no-cpu-baseline
r "textures" $B --workload textures --no-cpu-baseline
timeout 100 python tools/shard_probe.py > gpurun_out/${T}_shard_probe.jsonl 2>> gpurun_out/${T}_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${T}_launches_cornell_box.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/${T}_launches.log 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum
for W in cornell_box:12 obj_viewer:8 mandelbrot:8; do
  N=${W%%:*}; S=${W##*:}
  QZ_GRAPH=0 QZ_PIPELINES=1 timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/${T}_counters_${N}.csv \
      python tools/profile_step.py --workload $N --spp $S > gpurun_out/${T}_counters_${N}.log 2>&1
done
QZ_GRAPH=0 QZ_PIPELINES=1 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_step_flat|k_shade|k_sample|k_bin|k_albedo" --launch-skip 32 --launch-count 8 \
    -o gpurun_out/${T}_cornell_steady -f python tools/profile_step.py --workload cornell_box --spp 64 > gpurun_out/${T}_ncu_cornell.log 2>&1
QZ_GRAPH=0 QZ_PIPELINES=1 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:"k_memo" --launch-count 2 \
    -o gpurun_out/${T}_cornell_memo -f python tools/profile_step.py --workload cornell_box --spp 64 > gpurun_out/${T}_ncu_memo.log 2>&1
for W in obj_viewer mandelbrot; do
QZ_GRAPH=0 QZ_PIPELINES=1 timeout 300 ncu --set full --clock-control none --import-source on --kernel-name regex:k_trace_lane --launch-skip 6 --launch-count 2 \
    -o gpurun_out/${T}_${W}_trace -f python tools/profile_step.py --workload $W --spp 48 > gpurun_out/${T}_ncu_${W}.log 2>&1
done
tail -3 gpurun_out/${T}_tests.log; cut -c1-200 gpurun_out/${T}_bench_default.log; cat gpurun_out/${T}_shard_probe.jsonl
