#!/bin/bash
# Warm-cache per-launch durations (ncu, no cache flush, no clock control) of one eager iteration.  usage: TAG dataset...
TAG=$1; shift
O=gpurun_out; mkdir -p $O
for D in "$@"; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 3000 --csv --log-file $O/warm_${TAG}_${D}.csv \
    python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --dataset $D > $O/warm_${TAG}_${D}.log 2>&1; echo "ncu $D rc=$?"
done
