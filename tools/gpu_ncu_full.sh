#!/bin/bash
# ncu --set full of selected kernels in one eager bench iteration.  usage: TAG DATASET REGEX SKIP COUNT [extra bench args]
TAG=$1; D=$2; RX=$3; SKIP=$4; CNT=$5; shift 5
O=gpurun_out; mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$RX" -s $SKIP -c $CNT -f -o $O/full_${TAG} \
  python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --dataset $D "$@" > $O/full_${TAG}.log 2>&1; echo "ncu rc=$?"
ncu -i $O/full_${TAG}.ncu-rep --page raw --csv > $O/full_${TAG}_raw.csv 2>/dev/null
ncu -i $O/full_${TAG}.ncu-rep --page source --csv > $O/full_${TAG}_src.csv 2>/dev/null
ls -la $O/full_${TAG}*
