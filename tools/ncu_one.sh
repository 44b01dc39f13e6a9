#!/bin/bash
# usage: tools/ncu_one.sh TAG KERNEL_REGEX SKIP COUNT [prof_frame args...]
set -u
TAG=$1; K=$2; S=$3; C=$4; shift 4
FRAME="python tools/prof_frame.py --frames 2 $*"
$FRAME > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $C -f -o gpurun_out/${TAG} $FRAME > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/${TAG}_ncu.log
