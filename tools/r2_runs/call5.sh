#!/bin/bash
# Pipeline-depth experiment: default library vs the deep variant (4 A stages in TMEM + 4 weight stages on 128-wide tiles).
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
DEEP=$PWD/distributed-gan_b200/mdgan_b200/libmdgan_b200_deep.so
timeout 100 python tools/conv_bench.py 1 > $O/r2c5_convbench_default.log 2>&1
MDGAN_B200_LIB=$DEEP timeout 100 python tools/conv_bench.py 1 > $O/r2c5_convbench_deep.log 2>&1
MDGAN_B200_LIB=$DEEP timeout 200 python -m pytest tests/test_kernels_gpu.py -q -k "conv" 2>&1 | tail -5 > $O/r2c5_pytest_deep.log
timeout 150 python bench.py --dataset CelebA --steps 30 --warmup 5 --no-cpu-baseline --no-shapes > $O/r2c5_bench_default.json 2> $O/r2c5_bench_default.err
MDGAN_B200_LIB=$DEEP timeout 150 python bench.py --dataset CelebA --steps 30 --warmup 5 --no-cpu-baseline --no-shapes > $O/r2c5_bench_deep.json 2> $O/r2c5_bench_deep.err
paste <(cut -c1-60 $O/r2c5_convbench_default.log) <(cut -c26-60 $O/r2c5_convbench_deep.log)
cat $O/r2c5_pytest_deep.log
cut -c1-250 $O/r2c5_bench_default.json; cut -c1-250 $O/r2c5_bench_deep.json
