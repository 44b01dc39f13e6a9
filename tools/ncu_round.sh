#!/bin/bash
# Run on the GPU box (under gpurun): plain run, launch list of bench.py, instruction counts of one frame, and one
# --set full capture of the traversal kernels of one frame.  Outputs under gpurun_out/.
set -u
TAG=${1:-r02}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$?"
FRAME="python tools/prof_frame.py --frames 2"
$FRAME > gpurun_out/${TAG}_frame_plain.log 2>&1 &&
ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:"k_primary|k_emit|k_shadow|k_shade|k_bounce|k_overflow|k_resolve|k_subsample|k_supersample" -c 80 --csv --log-file gpurun_out/${TAG}_frame_metrics.csv $FRAME > gpurun_out/${TAG}_ncu_frame.log 2>&1
echo "frame metrics rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_shadow|k_primary|k_bounce" -s 6 -c 6 -f -o gpurun_out/${TAG}_trav $FRAME > gpurun_out/${TAG}_ncu_trav.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/${TAG}_ncu_trav.log
