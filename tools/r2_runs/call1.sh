#!/bin/bash
# Round-2 GPU call: smoke gate, suite, race hunt across kernel variants, transposer-group timing, parity report.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
nvidia-smi -L | head -2
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c1_smoke.log 2>&1 || { tail -30 gpurun_out/r2c1_smoke.log; echo SMOKE FAILED; exit 1; }
tail -2 gpurun_out/r2c1_smoke.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > gpurun_out/r2c1_pytest.log
tail -3 gpurun_out/r2c1_pytest.log
for ds in CIFAR10 MNIST_DCGAN CelebA; do
  MDGAN_CONV_TA=0 MDGAN_WGRAD_TA=0 MDGAN_BN_FUSED_STATS=0 timeout 600 python tools/stress_conv.py --dataset $ds --n 128 --reps 500 --save /tmp/ref_$ds.pt > gpurun_out/r2c1_stress_${ds}_nota.log 2>&1
  for tg in 1 2 3; do
    MDGAN_BN_FUSED_STATS=0 MDGAN_CONV_TG=$tg MDGAN_WGRAD_TG=$(( tg > 2 ? 2 : tg )) timeout 600 python tools/stress_conv.py --dataset $ds --n 128 --reps 3000 --against /tmp/ref_$ds.pt > gpurun_out/r2c1_stress_${ds}_tg$tg.log 2>&1
  done
  timeout 600 python tools/stress_conv.py --dataset $ds --n 128 --reps 3000 > gpurun_out/r2c1_stress_${ds}_fused.log 2>&1
done
MDGAN_BN_FUSED_STATS=0 timeout 900 python tools/stress_conv.py --dataset CIFAR10 --n 128 --reps 20000 --against /tmp/ref_CIFAR10.pt > gpurun_out/r2c1_stress_CIFAR10_long.log 2>&1
grep -h "stress_conv" gpurun_out/r2c1_stress_*.log | grep -v "first rep"
for tg in 1 2 3; do
  MDGAN_CONV_TG=$tg MDGAN_WGRAD_TG=$(( tg > 2 ? 2 : tg )) timeout 300 python tools/conv_bench.py 1 > gpurun_out/r2c1_convbench_tg$tg.log 2>&1
  MDGAN_CONV_TG=$tg MDGAN_WGRAD_TG=$(( tg > 2 ? 2 : tg )) timeout 600 python bench.py --dataset CelebA --steps 30 --warmup 5 --no-cpu-baseline --no-shapes > gpurun_out/r2c1_bench_celeba_tg$tg.json 2> gpurun_out/r2c1_bench_celeba_tg$tg.err
done
MDGAN_BN_FUSED_STATS=0 timeout 600 python bench.py --dataset CelebA --steps 30 --warmup 5 --no-cpu-baseline --no-shapes > gpurun_out/r2c1_bench_celeba_unfused.json 2> gpurun_out/r2c1_bench_celeba_unfused.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2c1_bench_default.json 2> gpurun_out/r2c1_bench_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2c1_bench_reference.json 2> gpurun_out/r2c1_bench_reference.err
timeout 1500 python tools/parity_report.py > gpurun_out/r2c1_parity_report.jsonl 2> gpurun_out/r2c1_parity_report.err
tail -3 gpurun_out/r2c1_parity_report.err
