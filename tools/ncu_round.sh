#!/bin/bash
# Run on the GPU box (under gpurun): plain run, launch list of bench.py, and one --set full capture of the
# traversal kernels of one frame.  Outputs under gpurun_out/.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $BENCH > gpurun_out/${TAG}_ncu_bench.log 2>&1
echo "launch list rc=$?"
FRAME="python tools/prof_frame.py --frames 2"
$FRAME > gpurun_out/${TAG}_frame_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_shadow|k_primary|k_bounce|k_overflow" -s 13 -c 12 -f -o gpurun_out/${TAG}_trav $FRAME > gpurun_out/${TAG}_ncu_trav.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/${TAG}_ncu_trav.log
