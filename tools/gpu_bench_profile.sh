#!/bin/bash
# One-GPU measurement pass: bench lines (MNIST / CIFAR / CelebA shapes), reference arm, ncu launch lists.
# usage (under gpurun): bash tools/gpu_bench_profile.sh TAG
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --steps 100 --warmup 10 > $O/bench_${TAG}_mnist.json 2> $O/bench_${TAG}_mnist.err
timeout 600 python bench.py --steps 50 --warmup 5 --dataset CIFAR10 --no-cpu-baseline > $O/bench_${TAG}_cifar.json 2> $O/bench_${TAG}_cifar.err
timeout 600 python bench.py --steps 50 --warmup 5 --dataset CelebA --no-cpu-baseline > $O/bench_${TAG}_celeba.json 2> $O/bench_${TAG}_celeba.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/launches_${TAG}_mnist.csv \
  python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline > $O/ncu_${TAG}_mnist.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_${TAG}_celeba.csv \
  python bench.py --steps 3 --warmup 3 --no-graph --no-cpu-baseline --dataset CelebA > $O/ncu_${TAG}_celeba.log 2>&1
echo done
