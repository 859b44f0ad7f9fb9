#!/bin/bash
# elect.sync MMA/TMA issue roles: correctness + timing.
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
timeout 180 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2c6_smoke.log 2>&1 || { tail -30 $O/r2c6_smoke.log; echo SMOKE FAILED; exit 1; }
tail -1 $O/r2c6_smoke.log | cut -c1-300
timeout 100 python tools/conv_bench.py 1 > $O/r2c6_convbench.log 2>&1; cat $O/r2c6_convbench.log
( timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > $O/r2c6_pytest.log; tail -3 $O/r2c6_pytest.log
timeout 150 python bench.py --dataset CelebA --steps 30 --warmup 5 --no-cpu-baseline > $O/r2c6_bench_celeba.json 2> $O/r2c6_bench_celeba.err; cut -c1-260 $O/r2c6_bench_celeba.json
MDGAN_CONV_TA=0 MDGAN_WGRAD_TA=0 MDGAN_BN_FUSED_STATS=0 timeout 100 python tools/stress_conv.py --dataset CIFAR10 --n 128 --reps 300 --save /tmp/ref.pt > $O/r2c6_stress_nota.log 2>&1
MDGAN_BN_FUSED_STATS=0 timeout 200 python tools/stress_conv.py --dataset CIFAR10 --n 128 --reps 20000 --against /tmp/ref.pt > $O/r2c6_stress_ta.log 2>&1
timeout 200 python tools/stress_conv.py --dataset CelebA --n 128 --reps 10000 > $O/r2c6_stress_celeba.log 2>&1
grep -h "stress_conv" $O/r2c6_stress_*.log | grep -v "first rep"
