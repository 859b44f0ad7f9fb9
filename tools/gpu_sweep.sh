#!/bin/bash
# BASELINE config 5: per-worker batch sweep at K = 1 (one GPU).  usage: TAG [precision]
TAG=$1; PREC=${2:-tf32x3}
O=gpurun_out; mkdir -p $O
for D in CIFAR10 CelebA; do
  for B in 32 64 128 256 512 1024; do
    timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dataset $D --batch $B --precision $PREC \
      > $O/sweep_${TAG}_${D}_b$B.json 2> $O/sweep_${TAG}_${D}_b$B.err; echo "$D b=$B rc=$?"
  done
done
